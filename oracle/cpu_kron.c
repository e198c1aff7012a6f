/* CPU implementation (C, OpenMP) of the SAME Kronecker-factored, cell-centric algorithm the CUDA kernels use, for fully
 * periodic 3-D boxes.  TEST / MEASUREMENT INFRASTRUCTURE, not product: bench.py reports it beside the oracle's C port as
 * the strongest CPU number we can produce on the box's host cores ("our algorithm on the CPU"), so that the GPU/CPU ratio
 * is not inflated by the port's quadrature-point evaluation.  It is NOT the reference's algorithm (that is cpu_vmult.c).
 *
 *   A u = vol (M (x) M (x) M) [ kappa u + sum_d ( Kt_d u + sum_s a_s FV_s + b_s FG_s ) ]          (DESIGN.md "Kronecker form")
 * with FV = -1/2 (G + G') + C (V - V'), FG = -1/2 (V - V')   (NS.cc:1284-1287).
 */
#include <omp.h>
#include <string.h>

#define MAXN 9

typedef struct {
  int n;
  double Kt[3][MAXN * MAXN]; /* (1/h_d^2) M^-1 K */
  double a[3][2][MAXN], b[3][2][MAXN];
  double g[2][MAXN], gn[2][MAXN];
  double Mh[3][MAXN * MAXN]; /* h_d M */
  double kappa, C;
} kron_params;

static inline void line_geom(int n, int d, int l, int *base, int *stride) {
  if (d == 0) { *base = l * n; *stride = 1; }
  else if (d == 1) { *base = (l % n) + n * n * (l / n); *stride = n; }
  else { *base = l; *stride = n * n; }
}

/* nbr[cell*6 + f]: all faces interior (periodic) */
void cpu_kron_vmult(const kron_params *P, long long n_cells, const long long *nbr, const double *src, double *dst) {
  const int n = P->n, nl = n * n, nd = nl * n;
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < n_cells; ++c) {
    const double *u = src + c * nd;
    double acc[MAXN * MAXN * MAXN], tmp[MAXN * MAXN * MAXN];
    for (int i = 0; i < nd; ++i) acc[i] = P->kappa * u[i];
    for (int d = 0; d < 3; ++d) {
      const double *un[2] = {src + nbr[c * 6 + 2 * d] * nd, src + nbr[c * 6 + 2 * d + 1] * nd};
      for (int l = 0; l < nl; ++l) {
        int base, stride;
        line_geom(n, d, l, &base, &stride);
        double x[MAXN], t[MAXN];
        for (int i = 0; i < n; ++i) x[i] = u[base + stride * i];
        for (int i = 0; i < n; ++i) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += P->Kt[d][i * n + j] * x[j];
          t[i] = s;
        }
        for (int s = 0; s < 2; ++s) {
          double G = 0, Gn = 0;
          for (int j = 0; j < n; ++j) { G += P->g[s][j] * x[j]; Gn += P->gn[s][j] * un[s][base + stride * j]; }
          const double V = x[s * (n - 1)], Vn = un[s][base + stride * ((1 - s) * (n - 1))];
          const double FV = -0.5 * (G + Gn) + P->C * (V - Vn), FG = -0.5 * (V - Vn);
          for (int i = 0; i < n; ++i) t[i] += P->a[d][s][i] * FV + P->b[d][s][i] * FG;
        }
        for (int i = 0; i < n; ++i) acc[base + stride * i] += t[i];
      }
    }
    for (int d = 0; d < 3; ++d) {
      for (int l = 0; l < nl; ++l) {
        int base, stride;
        line_geom(n, d, l, &base, &stride);
        double x[MAXN];
        for (int i = 0; i < n; ++i) x[i] = acc[base + stride * i];
        for (int i = 0; i < n; ++i) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += P->Mh[d][i * n + j] * x[j];
          tmp[base + stride * i] = s;
        }
      }
      memcpy(acc, tmp, sizeof(double) * nd);
    }
    memcpy(dst + c * nd, acc, sizeof(double) * nd);
  }
}

int cpu_kron_num_threads(void) { return omp_get_max_threads(); }
