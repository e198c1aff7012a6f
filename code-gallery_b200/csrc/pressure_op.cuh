// Shared declarations of the pressure (SIPG Helmholtz) operator kernels.
#pragma once
#include "dgx_internal.h"

namespace dgx {

__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

// Level-scaled 1-D tables of the Kronecker-factored operator (DESIGN.md "Kronecker form"):
//   A u = vol * (M (x) M (x) M) [ kappa u + sum_d  T_d u ],
//   T_d u|line = Kt_d u + sum_{s in faces} ( a_d,s FV_s + b_d,s FG_s )
// with the face fluxes of NS.cc:1284-1287 / 1320-1321 in reference scaling
//   interior:  FV = -1/2 (G + G') + C (V - V'),  FG = -1/2 (V - V')
//   id == 1:   FV = -G + C V,                    FG = -V           (weak Dirichlet)
//   other ids: FV = FG = 0                                          (homogeneous Neumann)
// V, G = own value / outward normal derivative trace, V', G' = the neighbour's.
template <typename T, int n>
struct PresParams {
  T Kt[MAX_DIM][n * n];   // (1/h_d^2) M^-1 K
  T a[MAX_DIM][2][n];     // (1/h_d^2) M^-1 f_s
  T b[MAX_DIM][2][n];     // (1/h_d^2) n_s M^-1 g_s
  T g[2][n];              // n_s g_s      : own outward normal derivative (reference scaling)
  T gn[2][n];             // n_s g_{1-s}  : neighbour's derivative along our outward normal
  T Mh[MAX_DIM][n * n];   // h_d M
  T kappa;                // 1/coeff, NS.cc:1237, 1248
  T C;                    // C_p = (p+1)^2, NS.cc:292
};

template <typename T, int n>
void fill_pres_params(PresParams<T, n> &P, const dgx_op *op, double kappa) {
  const Shape1D &s = get_shape(n - 1, n);
  const int dim = op->mf->dim;
  for (int d = 0; d < MAX_DIM; ++d) {
    const double h = d < dim ? op->mf->h[d] : 1.0, ih2 = 1.0 / (h * h);
    for (int i = 0; i < n * n; ++i) { P.Kt[d][i] = (T)(ih2 * s.Kt[i]); P.Mh[d][i] = (T)(h * s.M[i]); }
    for (int sd = 0; sd < 2; ++sd) {
      const double ns = sd ? 1.0 : -1.0;
      for (int i = 0; i < n; ++i) { P.a[d][sd][i] = (T)(ih2 * s.a[sd][i]); P.b[d][sd][i] = (T)(ih2 * ns * s.b[sd][i]); }
    }
  }
  for (int sd = 0; sd < 2; ++sd) {
    const double ns = sd ? 1.0 : -1.0;
    for (int i = 0; i < n; ++i) { P.g[sd][i] = (T)(ns * s.g[sd][i]); P.gn[sd][i] = (T)(ns * s.g[1 - sd][i]); }
  }
  P.kappa = (T)kappa;
  P.C = (T)((double)n * n);
}

// Optional fused epilogue of the operator kernels: instead of dst = A x they write
//   dst = a * x + b * xold + c * d .* (rhs - A x)        (null pointers drop their term; d == null means d = 1)
// which is one Chebyshev step (PreconditionChebyshev) or the multigrid residual, without the extra
// passes over t = A x.
template <typename T>
struct Epilogue {
  const T *xold = nullptr, *d = nullptr, *rhs = nullptr;
  T a = 0, b = 0, c = 0;
  int on = 0;
};

// one x-row of a cell (n contiguous values) from / to global memory, in 16-byte pieces when the row length allows it
template <typename T, int n>
__device__ __forceinline__ void load_row(const T *__restrict__ p, T (&u)[n]) {
  if constexpr (sizeof(T) == 8 && n % 2 == 0) {
#pragma unroll
    for (int k = 0; k < n / 2; ++k) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + k); u[2 * k] = t.x; u[2 * k + 1] = t.y; }
  } else if constexpr (sizeof(T) == 4 && n % 4 == 0) {
#pragma unroll
    for (int k = 0; k < n / 4; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + k);
      u[4 * k] = t.x; u[4 * k + 1] = t.y; u[4 * k + 2] = t.z; u[4 * k + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) u[i] = p[i];
  }
}
template <typename T, int n>
__device__ __forceinline__ void store_row(T *p, const T (&u)[n], int add) {
  if constexpr (sizeof(T) == 8 && n % 2 == 0) {
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      double2 t = make_double2(u[2 * k], u[2 * k + 1]);
      if (add) { const double2 o = reinterpret_cast<double2 *>(p)[k]; t.x += o.x; t.y += o.y; }
      reinterpret_cast<double2 *>(p)[k] = t;
    }
  } else if constexpr (sizeof(T) == 4 && n % 4 == 0) {
#pragma unroll
    for (int k = 0; k < n / 4; ++k) {
      float4 t = make_float4(u[4 * k], u[4 * k + 1], u[4 * k + 2], u[4 * k + 3]);
      if (add) { const float4 o = reinterpret_cast<float4 *>(p)[k]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
      reinterpret_cast<float4 *>(p)[k] = t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < n; ++i) p[i] = add ? p[i] + u[i] : u[i];
  }
}

double pressure_kappa(const dgx_op *op);
int pressure_apply_fast_fused(dgx_op *op, dgx_vec *dst, const dgx_vec *x, const FusedUpdate &fu);

// dst[cells] = A src for the owned cells [cell_begin, cell_end) only, on `stream` (generic kernel; the pipelined host entry
// point computes chunk by chunk while later chunks are still in flight over PCIe)
int pressure_apply_range(dgx_op *op, dgx_vec *dst, const dgx_vec *src, long long cell_begin, long long cell_end, cudaStream_t stream);

// blocked fast path (pressure_fast.cu): returns DGX_ERR_UNSUPPORTED when the shape does not qualify
int pressure_apply_fast(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add);
// multi-rank form: starts the ghost import of src itself, runs the tile columns that need no ghosts meanwhile
int pressure_apply_fast_overlapped(dgx_op *op, dgx_vec *dst, dgx_vec *src, bool add);

}  // namespace dgx
