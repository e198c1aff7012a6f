/* Vectorised CPU baseline of the pressure operator -- TEST / MEASUREMENT INFRASTRUCTURE, not product.
 *
 * The reference algorithm (NS.cc:1230-1292 inside MatrixFree::loop, NS.cc:1398-1421) executed the way deal.II executes it
 * on AVX hardware, restated in C because deal.II itself cannot be built in this image:
 *   - SIMD over cells: VL cells (4 with AVX2, 8 with AVX-512) per batch, one cell per lane (VectorizedArray<double>);
 *     dof values are transposed into struct-of-arrays on load (vectorized_load_and_transpose) and back on store;
 *   - compile-time degree (N = p + 1 = n_q), fully unrolled 1-D kernels with the even-odd decomposition
 *     (EvaluatorTensorProduct<evaluate_evenodd>): a 5 x 5 line costs 13 FMA + 8 add instead of 25 FMA;
 *   - cell term in the collocation form deal.II picks for n_q = n: basis change to the Gauss points (3 sweeps), collocation
 *     derivative (3), quadrature-point loop, transposed derivative (3), transposed basis change (3);
 *   - face term face-centric: every interior face is visited once from its lower cell, both sides are evaluated
 *     (normal contraction, 2 tangential sweeps per quantity), the SIPG fluxes of NS.cc:1284-1287 are formed at the face
 *     quadrature points, both sides are integrated and added to dst (dst is zeroed first, as Base::vmult does);
 *   - one thread per core over contiguous chunks of the cell order stands in for one MPI rank per core (NS.cc:2304
 *     disables TBB); a face between two chunks is evaluated by both owners, each adding only its own side (deal.II
 *     evaluates it once and communicates instead: ghost import + compress).
 * Periodic 3-D boxes only (the BASELINE workload); boundary faces are not handled here (cpu_vmult.c does those).
 */
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#ifndef VL
#define VL 8
#endif
#ifndef N
#define N 5
#endif
#define NH (N / 2)
#define NM ((N + 1) / 2) /* rows of the even-odd tables: q < NH plus the middle one for odd N */
#define NL (N * N)
#define ND (N * N * N)

typedef double vd __attribute__((vector_size(VL * 8)));

typedef struct {
  double Ae[NM][NH > 0 ? NH : 1], Ao[NM][NH > 0 ? NH : 1], Am[NM];
  int sgn; /* +1: A[n-1-q][n-1-i] = A[q][i]   -1: = -A[q][i] */
} eo_mat;

typedef struct {
  int n;
  double S[N * N];    /* S[q][i]: nodal basis at the Gauss points */
  double Dh[N * N];   /* Dh[q][r]: derivative of the Gauss-point Lagrange basis at the Gauss points */
  double w[N];
  double f[2][N], g[2][N];   /* l_i(s), l_i'(s) at the two end points */
  double h[3];
  double kappa, C;
} simd_params;

typedef struct {
  eo_mat S, St, D, Dt;
  double w[N], f[2][N], g[2][N], h[3], kappa, C;
  int nodal_faces; /* f[s] is a unit vector: the value trace is a pick */
} simd_tables;

static void eo_setup(eo_mat *m, const double *A, int transpose, int sgn) {
  m->sgn = sgn;
  for (int q = 0; q < NM; ++q) {
    for (int i = 0; i < NH; ++i) {
      const double a = transpose ? A[i * N + q] : A[q * N + i], b = transpose ? A[(N - 1 - i) * N + q] : A[q * N + (N - 1 - i)];
      m->Ae[q][i] = 0.5 * (a + b);
      m->Ao[q][i] = 0.5 * (a - b);
    }
    m->Am[q] = (N & 1) ? (transpose ? A[NH * N + q] : A[q * N + NH]) : 0.0;
  }
}

/* one line: out = A in, even-odd form */
static inline __attribute__((always_inline)) void eo_line(const eo_mat *m, const int sgn, const vd *in, int stride, vd *out, int ostride) {
  vd e[NH > 0 ? NH : 1], o[NH > 0 ? NH : 1], x[N];
  for (int i = 0; i < N; ++i) x[i] = in[i * stride];
  for (int i = 0; i < NH; ++i) { e[i] = x[i] + x[N - 1 - i]; o[i] = x[i] - x[N - 1 - i]; }
  for (int q = 0; q < NH; ++q) {
    vd E = m->Ae[q][0] * e[0], O = m->Ao[q][0] * o[0];
    for (int i = 1; i < NH; ++i) { E += m->Ae[q][i] * e[i]; O += m->Ao[q][i] * o[i]; }
    if (N & 1) E += m->Am[q] * x[NH];
    out[q * ostride] = E + O;
    out[(N - 1 - q) * ostride] = sgn > 0 ? E - O : O - E;
  }
  if (N & 1) {
    vd r;
    if (sgn > 0) { r = m->Am[NH] * x[NH]; for (int i = 0; i < NH; ++i) r += m->Ae[NH][i] * e[i]; }
    else { r = m->Ao[NH][0] * o[0]; for (int i = 1; i < NH; ++i) r += m->Ao[NH][i] * o[i]; }
    out[NH * ostride] = r;
  }
}

/* sweep along direction d of an N^3 block (in place) */
static inline __attribute__((always_inline)) void sweep3(const eo_mat *m, const int sgn, vd *a, const int d) {
  const int stride = d == 0 ? 1 : (d == 1 ? N : NL);
  for (int l = 0; l < NL; ++l) {
    const int base = d == 0 ? l * N : (d == 1 ? (l % N) + NL * (l / N) : l);
    eo_line(m, sgn, a + base, stride, a + base, stride);
  }
}
/* sweep along direction e of an N^2 face array (in place) */
static inline __attribute__((always_inline)) void sweep2(const eo_mat *m, vd *a, const int e) {
  const int stride = e == 0 ? 1 : N;
  for (int l = 0; l < N; ++l) {
    const int base = e == 0 ? l * N : l;
    eo_line(m, +1, a + base, stride, a + base, stride);
  }
}

/* NS.cc:1230-1252: v_i += int grad phi_i . grad p + kappa phi_i p */
static void cell_term(const simd_tables *T, const vd *u, vd *out) {
  vd a[ND], gr[ND], acc[ND];
  memcpy(a, u, sizeof(a));
  sweep3(&T->S, +1, a, 0); sweep3(&T->S, +1, a, 1); sweep3(&T->S, +1, a, 2);   /* values at the Gauss points */
  const double vol = T->h[0] * T->h[1] * T->h[2];
  for (int q = 0; q < ND; ++q) {
    const double JxW = vol * T->w[q % N] * T->w[(q / N) % N] * T->w[q / NL];
    acc[q] = (T->kappa * JxW) * a[q];   /* submit_value */
  }
  for (int d = 0; d < 3; ++d) {
    const int stride = d == 0 ? 1 : (d == 1 ? N : NL);
    const double ih2 = 1.0 / (T->h[d] * T->h[d]);
    for (int l = 0; l < NL; ++l) {      /* collocation derivative, scaled by J^-T J^-1 JxW (submit_gradient), and its transpose */
      const int base = d == 0 ? l * N : (d == 1 ? (l % N) + NL * (l / N) : l);
      vd t[N], r[N];
      eo_line(&T->D, -1, a + base, stride, t, 1);
      for (int i = 0; i < N; ++i) {
        const int q = base + i * stride;
        t[i] *= ih2 * vol * T->w[q % N] * T->w[(q / N) % N] * T->w[q / NL];
      }
      eo_line(&T->Dt, -1, t, 1, r, 1);
      for (int i = 0; i < N; ++i) gr[base + i * stride] = r[i];
    }
    for (int q = 0; q < ND; ++q) acc[q] += gr[q];
  }
  sweep3(&T->St, +1, acc, 0); sweep3(&T->St, +1, acc, 1); sweep3(&T->St, +1, acc, 2);   /* test with the nodal basis */
  for (int i = 0; i < ND; ++i) out[i] = acc[i];
}

/* value and normal derivative (d/dx_d) of u on face (d, side) at the face quadrature points */
static inline void face_eval(const simd_tables *T, const vd *u, int d, int side, vd *val, vd *dn) {
  const int stride = d == 0 ? 1 : (d == 1 ? N : NL);
  const double ih = 1.0 / T->h[d];
  for (int l = 0; l < NL; ++l) {
    const int base = d == 0 ? l * N : (d == 1 ? (l % N) + NL * (l / N) : l);
    vd v, g = (T->g[side][0] * ih) * u[base];
    for (int i = 1; i < N; ++i) g += (T->g[side][i] * ih) * u[base + i * stride];
    if (T->nodal_faces) v = u[base + (side ? N - 1 : 0) * stride];
    else { v = T->f[side][0] * u[base]; for (int i = 1; i < N; ++i) v += T->f[side][i] * u[base + i * stride]; }
    val[l] = v; dn[l] = g;
  }
  for (int e = 0; e < 2; ++e) { sweep2(&T->S, val, e); sweep2(&T->S, dn, e); }
}

/* out += phi-test of fv and d(phi)/dx_d-test of fg on face (d, side) */
static inline void face_integrate(const simd_tables *T, vd *out, int d, int side, vd *fv, vd *fg) {
  const int stride = d == 0 ? 1 : (d == 1 ? N : NL);
  const double ih = 1.0 / T->h[d];
  for (int e = 0; e < 2; ++e) { sweep2(&T->St, fv, e); sweep2(&T->St, fg, e); }
  for (int l = 0; l < NL; ++l) {
    const int base = d == 0 ? l * N : (d == 1 ? (l % N) + NL * (l / N) : l);
    if (T->nodal_faces) {
      for (int i = 0; i < N; ++i) out[base + i * stride] += (T->g[side][i] * ih) * fg[l];
      out[base + (side ? N - 1 : 0) * stride] += fv[l];
    } else {
      for (int i = 0; i < N; ++i) out[base + i * stride] += T->f[side][i] * fv[l] + (T->g[side][i] * ih) * fg[l];
    }
  }
}

/* NS.cc:1259-1292: interior face between um (its upper face in d) and up (its lower face); om / op may be NULL */
static void face_term(const simd_tables *T, const vd *um, const vd *up, vd *om, vd *op, int d) {
  vd vm[NL], gm[NL], vp[NL], gp[NL], sv[NL], sg[NL];
  face_eval(T, um, d, 1, vm, gm);
  face_eval(T, up, d, 0, vp, gp);
  const double sigma = T->C * 0.5 * (1.0 / T->h[d] + 1.0 / T->h[d]);
  const double area = T->h[0] * T->h[1] * T->h[2] / T->h[d];
  for (int q = 0; q < NL; ++q) {
    const double JxW = area * T->w[q % N] * T->w[q / N];
    const vd avg = 0.5 * (gm[q] + gp[q]), jump = vm[q] - vp[q];
    sv[q] = (sigma * jump - avg) * JxW;   /* submit_value on the interior side; the exterior side gets the negative */
    sg[q] = (-0.5 * JxW) * jump;          /* submit_normal_derivative, same on both sides */
  }
  if (om) {
    vd a[NL], b[NL];
    memcpy(a, sv, sizeof(a)); memcpy(b, sg, sizeof(b));
    face_integrate(T, om, d, 1, a, b);
  }
  if (op) {
    for (int q = 0; q < NL; ++q) sv[q] = -sv[q];
    face_integrate(T, op, d, 0, sv, sg);
  }
}

/* vectorized_load_and_transpose: lane l of u[i] = dof i of cell cells[l] */
static inline void load_batch(const double *src, const long long *cells, vd *u) {
  double *b = (double *)u;
  for (int l = 0; l < VL; ++l) {
    const double *p = src + cells[l] * ND;
    for (int i = 0; i < ND; ++i) b[i * VL + l] = p[i];
  }
}

void *simd_setup(const simd_params *P) {
  simd_tables *T = (simd_tables *)aligned_alloc(64, (sizeof(simd_tables) + 63) / 64 * 64);
  eo_setup(&T->S, P->S, 0, +1); eo_setup(&T->St, P->S, 1, +1);
  eo_setup(&T->D, P->Dh, 0, -1); eo_setup(&T->Dt, P->Dh, 1, -1);
  memcpy(T->w, P->w, sizeof(T->w)); memcpy(T->f, P->f, sizeof(T->f)); memcpy(T->g, P->g, sizeof(T->g)); memcpy(T->h, P->h, sizeof(T->h));
  T->kappa = P->kappa; T->C = P->C;
  T->nodal_faces = 1;
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < N; ++i) {
      const double want = (i == (s ? N - 1 : 0)) ? 1.0 : 0.0, dev = P->f[s][i] - want;
      if (dev > 1e-14 || dev < -1e-14) T->nodal_faces = 0;
    }
  return T;
}
void simd_free(void *T) { free(T); }
int simd_vl(void) { return VL; }
int simd_n(void) { return N; }
int simd_num_threads(void) { return omp_get_max_threads(); }

/* nbr[cell * 6 + f], all faces interior (periodic box), n_cells a multiple of VL.  dst = A src */
void simd_pressure_vmult(const void *tables, long long n_cells, const long long *nbr, const double *src, double *dst, int n_threads) {
  const simd_tables *T = (const simd_tables *)tables;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
  {
    const int nt = omp_get_num_threads(), me = omp_get_thread_num();
    const long long nb = n_cells / VL;
    const long long c_begin = (nb * me / nt) * VL, c_end = (nb * (me + 1) / nt) * VL;
    memset(dst + c_begin * ND, 0, sizeof(double) * (size_t)(c_end - c_begin) * ND);   /* Base::vmult: dst = 0 */
    vd u[ND], un[ND], acc[ND], oth[ND];
    long long own[VL], nbc[VL];
    for (long long c0 = c_begin; c0 < c_end; c0 += VL) {
      for (int l = 0; l < VL; ++l) own[l] = c0 + l;
      load_batch(src, own, u);
      cell_term(T, u, acc);
      for (int d = 0; d < 3; ++d) {
        /* upper face: visited from here; the neighbour's share is added to dst when this thread owns the neighbour */
        int all_mine = 1, any_mine = 0;
        for (int l = 0; l < VL; ++l) {
          nbc[l] = nbr[(c0 + l) * 6 + 2 * d + 1];
          const int mine = nbc[l] >= c_begin && nbc[l] < c_end;
          all_mine &= mine; any_mine |= mine;
        }
        load_batch(src, nbc, un);
        if (any_mine) memset(oth, 0, sizeof(oth));
        face_term(T, u, un, acc, any_mine ? oth : NULL, d);
        if (any_mine)
          for (int l = 0; l < VL; ++l) {
            if (!(nbc[l] >= c_begin && nbc[l] < c_end)) continue;
            double *p = dst + nbc[l] * ND;
            const double *b = (const double *)oth;
            for (int i = 0; i < ND; ++i) p[i] += b[i * VL + l];
          }
        /* lower face: only where the lower neighbour belongs to another thread (that thread adds its own side only) */
        int any_foreign = 0;
        for (int l = 0; l < VL; ++l) {
          nbc[l] = nbr[(c0 + l) * 6 + 2 * d];
          any_foreign |= !(nbc[l] >= c_begin && nbc[l] < c_end);
        }
        if (any_foreign) {
          load_batch(src, nbc, un);
          memset(oth, 0, sizeof(oth));
          face_term(T, un, u, NULL, oth, d);
          for (int l = 0; l < VL; ++l) {
            if (nbc[l] >= c_begin && nbc[l] < c_end) continue;
            for (int i = 0; i < ND; ++i) ((double *)acc)[i * VL + l] += ((const double *)oth)[i * VL + l];
          }
        }
        (void)all_mine;
      }
      for (int l = 0; l < VL; ++l) {
        double *p = dst + (c0 + l) * ND;
        const double *b = (const double *)acc;
        for (int i = 0; i < ND; ++i) p[i] += b[i * VL + l];
      }
    }
  }
}
