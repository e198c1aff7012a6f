/* CPU restatement (C, OpenMP) of the reference algorithm for the pressure operator, used ONLY as the
 * timed CPU baseline of bench.py and as a cross-check of the numpy oracle (tests/).  TEST
 * INFRASTRUCTURE, not product.  "CPU restatement of the reference algorithm, not deal.II".
 *
 * Structure follows what deal.II executes for NS.cc:1398-1421 with NS_stage == 2:
 *   cell loop  (NS.cc:1230-1252): sum-factorised evaluate (values + gradients at the Gauss points),
 *              quadrature-point loop, sum-factorised integrate;
 *   face loop  (NS.cc:1259-1292): face-centric, both sides evaluated, fluxes, both sides integrated;
 *   boundary   (NS.cc:1299-1326): boundary_id == 1 only.
 * Threads stand in for the reference's one-MPI-rank-per-core layout (NS.cc:2304 disables TBB);
 * faces are coloured by parity of the lower cell's coordinate so that threads never share a cell.
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 9

typedef struct {
  int dim, n;            /* n = p + 1 = n_q */
  double S[MAXN * MAXN]; /* S[q][i] */
  double D[MAXN * MAXN];
  double w[MAXN];
  double f[2][MAXN], g[2][MAXN];
  double h[3];
  double kappa, C;
} cpu_params;

static inline void apply_line(const double *M, int n, int transpose, const double *in, int stride, double *out) {
  for (int q = 0; q < n; ++q) {
    double acc = 0;
    for (int i = 0; i < n; ++i) acc += (transpose ? M[i * n + q] : M[q * n + i]) * in[i * stride];
    out[q * stride] = acc;
  }
}

/* apply M along direction d of an n^3 (or n^2) block, out may not alias in */
static void sweep(const double *M, int n, int dim, int d, int transpose, const double *in, double *out) {
  const int nd = dim == 3 ? n * n * n : n * n;
  int stride = 1;
  for (int e = 0; e < d; ++e) stride *= n;
  for (int base = 0; base < nd; ++base) {
    if ((base / stride) % n != 0) continue;
    apply_line(M, n, transpose, in + base, stride, out + base);
  }
}

static void cell_term(const cpu_params *P, const double *u, double *out) {
  const int n = P->n, dim = P->dim, nd = dim == 3 ? n * n * n : n * n;
  double a[MAXN * MAXN * MAXN], b[MAXN * MAXN * MAXN], val[MAXN * MAXN * MAXN], acc[MAXN * MAXN * MAXN];
  double grad[3][MAXN * MAXN * MAXN];
  /* values at quadrature points */
  memcpy(a, u, sizeof(double) * nd);
  for (int d = 0; d < dim; ++d) { sweep(P->S, n, dim, d, 0, a, b); memcpy(a, b, sizeof(double) * nd); }
  memcpy(val, a, sizeof(double) * nd);
  /* gradients: D in direction e, S in the others */
  for (int e = 0; e < dim; ++e) {
    memcpy(a, u, sizeof(double) * nd);
    for (int d = 0; d < dim; ++d) { sweep(d == e ? P->D : P->S, n, dim, d, 0, a, b); memcpy(a, b, sizeof(double) * nd); }
    for (int q = 0; q < nd; ++q) grad[e][q] = a[q] / P->h[e];
  }
  /* quadrature loop: submit_gradient(grad), submit_value(kappa * val) */
  double vol = 1;
  for (int d = 0; d < dim; ++d) vol *= P->h[d];
  for (int q = 0; q < nd; ++q) {
    int qq = q;
    double JxW = vol;
    for (int d = 0; d < dim; ++d) { JxW *= P->w[qq % n]; qq /= n; }
    val[q] = P->kappa * val[q] * JxW;
    for (int e = 0; e < dim; ++e) grad[e][q] = grad[e][q] * JxW / P->h[e];
  }
  /* integrate */
  memcpy(a, val, sizeof(double) * nd);
  for (int d = 0; d < dim; ++d) { sweep(P->S, n, dim, d, 1, a, b); memcpy(a, b, sizeof(double) * nd); }
  memcpy(acc, a, sizeof(double) * nd);
  for (int e = 0; e < dim; ++e) {
    memcpy(a, grad[e], sizeof(double) * nd);
    for (int d = 0; d < dim; ++d) { sweep(d == e ? P->D : P->S, n, dim, d, 1, a, b); memcpy(a, b, sizeof(double) * nd); }
    for (int i = 0; i < nd; ++i) acc[i] += a[i];
  }
  for (int i = 0; i < nd; ++i) out[i] += acc[i];
}

/* traces on face (d, side): value and normal derivative (d/dx_d) at the face quadrature points */
static void face_eval(const cpu_params *P, const double *u, int d, int side, double *val, double *dn) {
  const int n = P->n, dim = P->dim, nf = dim == 3 ? n * n : n;
  int stride = 1;
  for (int e = 0; e < d; ++e) stride *= n;
  double tv[MAXN * MAXN], tg[MAXN * MAXN], tmp[MAXN * MAXN];
  /* contract direction d */
  int k = 0;
  const int nd = dim == 3 ? n * n * n : n * n;
  for (int base = 0; base < nd; ++base) {
    if ((base / stride) % n != 0) continue;
    double v = 0, g = 0;
    for (int i = 0; i < n; ++i) { v += P->f[side][i] * u[base + i * stride]; g += P->g[side][i] * u[base + i * stride]; }
    tv[k] = v; tg[k] = g / P->h[d]; ++k;
  }
  /* interpolate the remaining dim-1 directions to the Gauss points */
  for (int pass = 0; pass < 2; ++pass) {
    double *t = pass ? tg : tv;
    for (int e = 0; e < dim - 1; ++e) { sweep(P->S, n, dim - 1, e, 0, t, tmp); memcpy(t, tmp, sizeof(double) * nf); }
  }
  memcpy(val, tv, sizeof(double) * nf);
  memcpy(dn, tg, sizeof(double) * nf);
}

/* test with phi (fv) and d(phi)/dx_d (fg) on face (d, side); values already multiplied by JxW */
static void face_integrate(const cpu_params *P, double *out, int d, int side, double *fv, double *fg) {
  const int n = P->n, dim = P->dim, nf = dim == 3 ? n * n : n, nd = dim == 3 ? n * n * n : n * n;
  double tmp[MAXN * MAXN];
  for (int pass = 0; pass < 2; ++pass) {
    double *t = pass ? fg : fv;
    for (int e = 0; e < dim - 1; ++e) { sweep(P->S, n, dim - 1, e, 1, t, tmp); memcpy(t, tmp, sizeof(double) * nf); }
  }
  int stride = 1;
  for (int e = 0; e < d; ++e) stride *= n;
  int k = 0;
  for (int base = 0; base < nd; ++base) {
    if ((base / stride) % n != 0) continue;
    for (int i = 0; i < n; ++i) out[base + i * stride] += P->f[side][i] * fv[k] + P->g[side][i] / P->h[d] * fg[k];
    ++k;
  }
}

static void face_term(const cpu_params *P, const double *um, const double *up, double *om, double *op, int d) {
  const int n = P->n, dim = P->dim, nf = dim == 3 ? n * n : n;
  double vm[MAXN * MAXN], gm[MAXN * MAXN], vp[MAXN * MAXN], gp[MAXN * MAXN];
  double sv_m[MAXN * MAXN], sv_p[MAXN * MAXN], sg[MAXN * MAXN], sg2[MAXN * MAXN];
  face_eval(P, um, d, 1, vm, gm); /* interior side: upper face */
  face_eval(P, up, d, 0, vp, gp); /* exterior side: lower face */
  const double sigma = P->C * 0.5 * (1.0 / P->h[d] + 1.0 / P->h[d]);
  double area = 1;
  for (int e = 0; e < dim; ++e) if (e != d) area *= P->h[e];
  for (int q = 0; q < nf; ++q) {
    int qq = q;
    double JxW = area;
    for (int e = 0; e < dim - 1; ++e) { JxW *= P->w[qq % n]; qq /= n; }
    const double avg = 0.5 * (gm[q] + gp[q]), jump = vm[q] - vp[q];
    sv_m[q] = (-avg + sigma * jump) * JxW;
    sv_p[q] = (avg - sigma * jump) * JxW;
    sg[q] = -0.5 * jump * JxW;
    sg2[q] = sg[q];
  }
  face_integrate(P, om, d, 1, sv_m, sg);
  face_integrate(P, op, d, 0, sv_p, sg2);
}

static void boundary_term(const cpu_params *P, const double *u, double *out, int d, int side) {
  const int n = P->n, dim = P->dim, nf = dim == 3 ? n * n : n;
  double v[MAXN * MAXN], g[MAXN * MAXN], sv[MAXN * MAXN], sg[MAXN * MAXN];
  face_eval(P, u, d, side, v, g);
  const double ns = side ? 1.0 : -1.0, sigma = P->C / P->h[d];
  double area = 1;
  for (int e = 0; e < dim; ++e) if (e != d) area *= P->h[e];
  for (int q = 0; q < nf; ++q) {
    int qq = q;
    double JxW = area;
    for (int e = 0; e < dim - 1; ++e) { JxW *= P->w[qq % n]; qq /= n; }
    sv[q] = (-ns * g[q] + sigma * v[q]) * JxW;
    sg[q] = -v[q] * ns * JxW; /* submit_normal_derivative(-theta p): d/dn = ns d/dx_d */
  }
  face_integrate(P, out, d, side, sv, sg);
}

/* nbr[cell*2*dim + f]; coords[cell*dim + d]; dst = A src */
void cpu_pressure_vmult(const cpu_params *P, long long n_cells, const long long *nbr, const int *coords, const double *src, double *dst) {
  const int dim = P->dim, n = P->n, nd = dim == 3 ? n * n * n : n * n;
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < n_cells; ++c) {
    memset(dst + c * nd, 0, sizeof(double) * nd);
    cell_term(P, src + c * nd, dst + c * nd);
    for (int f = 0; f < 2 * dim; ++f)
      if (nbr[c * 2 * dim + f] == -2) boundary_term(P, src + c * nd, dst + c * nd, f / 2, f % 2);
  }
  for (int d = 0; d < dim; ++d)
    for (int colour = 0; colour < 3; ++colour) {
#pragma omp parallel for schedule(static)
      for (long long c = 0; c < n_cells; ++c) {
        const long long nb = nbr[c * 2 * dim + 2 * d + 1];
        if (nb < 0) continue;
        /* colours 0/1: parity of the lower cell; colour 2: wrap-around faces of odd-sized periodic rows */
        const int wrap = coords[nb * dim + d] < coords[c * dim + d];
        const int col = wrap ? 2 : (coords[c * dim + d] & 1);
        if (col != colour) continue;
        if (nb == c) {
          double tmp[MAXN * MAXN * MAXN];
          memset(tmp, 0, sizeof(double) * nd);
          face_term(P, src + c * nd, src + nb * nd, dst + c * nd, tmp, d);
          for (int i = 0; i < nd; ++i) dst[c * nd + i] += tmp[i];
        } else {
          face_term(P, src + c * nd, src + nb * nd, dst + c * nd, dst + nb * nd, d);
        }
      }
    }
}

int cpu_num_threads(void) { return omp_get_max_threads(); }
