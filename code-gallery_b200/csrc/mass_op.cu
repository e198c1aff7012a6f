// DG mass operator, NS_stage == 3: assemble_cell_term_projection_grad_p (NS.cc:1373-1390), applied by
// cell_loop (NS.cc:1415-1418).  With a tensor-product quadrature the cell mass matrix is exactly
// (h_x M) (x) (h_y M) (x) (h_z M) with M_ij = sum_q w_q l_i(xi_q) l_j(xi_q), per component.
#include <algorithm>

#include "pressure_op.cuh"
#include "reduce.cuh"

namespace dgx {
namespace {

template <typename T, int n>
struct MassParams {
  T Mh[MAX_DIM][n * n];
};

template <int dim, int n>
__host__ __device__ constexpr int mass_cpb() { return (192 / ipow(n, dim - 1)) > 0 ? (192 / ipow(n, dim - 1)) : 1; }

// one "item" = one (cell, component) block of n^dim values; contiguous in the DG layout
template <int dim, int n, typename T>
__global__ void __launch_bounds__(mass_cpb<dim, n>() * ipow(n, dim - 1))
mass_kernel(const __grid_constant__ MassParams<T, n> P, long long n_items, const T *__restrict__ src, T *__restrict__ dst, int add) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = mass_cpb<dim, n>(), NT = CPB * NL;
  __shared__ T sa[CPB * ND];
  const int tid = threadIdx.x;
  const long long item0 = (long long)blockIdx.x * CPB;
  const int nblk = (int)min((long long)CPB, n_items - item0);
  for (int idx = tid; idx < nblk * ND; idx += NT) sa[idx] = src[(size_t)item0 * ND + idx];
  __syncthreads();
  const int c = tid / NL, l = tid % NL;
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    if (c < nblk) {
      int base, stride;
      if (d == 0) { base = l * n; stride = 1; }
      else if (d == 1 && dim == 3) { base = (l % n) + n * n * (l / n); stride = n; }
      else if (d == 1) { base = l; stride = n; }
      else { base = l; stride = n * n; }
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = sa[c * ND + base + stride * i];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) acc += P.Mh[d][i * n + j] * u[j];
        sa[c * ND + base + stride * i] = acc;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nblk * ND; idx += NT) {
    const size_t o = (size_t)item0 * ND + idx;
    dst[o] = add ? dst[o] + sa[idx] : sa[idx];
  }
}

// Front half of one SolverCG iteration on the mass matrix (project_grad, NS.cc:2574-2576; identity preconditioner), one pass:
//   p = r + beta p   (first iteration: p = r),   v = M p,   *result = p . v
// The search direction is built while the cell is loaded, the dot product while the result is stored; per-block partial
// sums are combined in a fixed order by the last block (deterministic).  Persistent grid (<= 8 blocks per SM).

template <int dim, int n, typename T>
__global__ void __launch_bounds__(mass_cpb<dim, n>() * ipow(n, dim - 1))
mass_cg_front_kernel(const __grid_constant__ MassParams<T, n> P, long long n_items, T *__restrict__ p, const T *__restrict__ r, T beta, int first,
                     T *__restrict__ v, double *partial, double *result) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = mass_cpb<dim, n>(), NT = CPB * NL;
  __shared__ T sa[CPB * ND];
  __shared__ double red[(NT + 31) / 32];
  __shared__ bool last;
  const int tid = threadIdx.x;
  const int c = tid / NL, l = tid % NL;
  double acc = 0.0;
  for (long long item0 = (long long)blockIdx.x * CPB; item0 < n_items; item0 += (long long)gridDim.x * CPB) {
    const int nblk = (int)min((long long)CPB, n_items - item0);
    T pk[n], rk[n];   // this thread's entries of p (the same entries it stores below); all loads before the first store
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const int idx = tid + k * NT;
      const size_t o = (size_t)item0 * ND + idx;
      const bool ok = idx < nblk * ND;
      rk[k] = ok ? r[o] : T(0);
      pk[k] = (ok && !first) ? p[o] : T(0);
    }
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const int idx = tid + k * NT;
      if (idx < nblk * ND) {
        const T x = first ? rk[k] : beta * pk[k] + rk[k];
        p[(size_t)item0 * ND + idx] = x; pk[k] = x; sa[idx] = x;
      }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      if (c < nblk) {
        int base, stride;
        if (d == 0) { base = l * n; stride = 1; }
        else if (d == 1 && dim == 3) { base = (l % n) + n * n * (l / n); stride = n; }
        else if (d == 1) { base = l; stride = n; }
        else { base = l; stride = n * n; }
        T u[n];
#pragma unroll
        for (int i = 0; i < n; ++i) u[i] = sa[c * ND + base + stride * i];
#pragma unroll
        for (int i = 0; i < n; ++i) {
          T a2 = 0;
#pragma unroll
          for (int j = 0; j < n; ++j) a2 += P.Mh[d][i * n + j] * u[j];
          sa[c * ND + base + stride * i] = a2;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const int idx = tid + k * NT;
      if (idx < nblk * ND) {
        const T y = sa[idx];
        v[(size_t)item0 * ND + idx] = y;
        acc += (double)pk[k] * (double)y;
      }
    }
    __syncthreads();
  }
  acc = warp_sum_existing(acc);   // NT is not a multiple of 32 for most degrees
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double s2 = 0;
    for (int w = 0; w < (NT + 31) / 32; ++w) s2 += red[w];
    partial[blockIdx.x] = s2;
    __threadfence();
    last = (atomicInc(scratch_ticket(partial), gridDim.x - 1) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && tid < 32) {
    __threadfence();
    double s2 = 0;
    for (int b = tid; b < (int)gridDim.x; b += 32) s2 += ((volatile double *)partial)[b];
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_down_sync(0xffffffffu, s2, o);
    if (tid == 0) *result = s2;
  }
}

template <int dim, int n, typename T>
int launch_mass_cg_front(dgx_op *op, dgx_vec *p, const dgx_vec *r, double beta, bool first, dgx_vec *v, double *pv) {
  dgx_mf *mf = op->mf;
  dgx_ctx *ctx = mf->ctx;
  const Shape1D &s = get_shape(n - 1, n);
  MassParams<T, n> P;
  for (int d = 0; d < MAX_DIM; ++d) {
    const double h = d < dim ? mf->h[d] : 1.0;
    for (int i = 0; i < n * n; ++i) P.Mh[d][i] = (T)(h * s.M[i]);
  }
  const long long items = mf->n_owned * op->ncomp;
  constexpr int CPB = mass_cpb<dim, n>(), NT = CPB * ipow(n, dim - 1);
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + CPB - 1) / CPB, 148 * 8));
  double *partial = ctx->d_scratch + kScratchPartial, *result = ctx->d_scratch;
  mass_cg_front_kernel<dim, n, T><<<grid, NT, 0, ctx->stream>>>(P, items, (T *)p->d, (const T *)r->d, (T)beta, first ? 1 : 0, (T *)v->d, partial, result);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  int rc = allreduce_sum(ctx, result, 1);
  if (rc) return rc;
  DGX_CUDA_CHECK(cudaMemcpyAsync(ctx->h_scratch, result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *pv = ctx->h_scratch[0];
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_mass(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  dgx_mf *mf = op->mf;
  const Shape1D &s = get_shape(n - 1, n);
  MassParams<T, n> P;
  for (int d = 0; d < MAX_DIM; ++d) {
    const double h = d < dim ? mf->h[d] : 1.0;
    for (int i = 0; i < n * n; ++i) P.Mh[d][i] = (T)(inverse ? s.Minv[i] / h : h * s.M[i]);   // (h M)^-1 = M^-1 / h
  }
  const long long items = mf->n_owned * op->ncomp;
  if (items == 0) return DGX_OK;
  constexpr int CPB = mass_cpb<dim, n>(), NT = CPB * ipow(n, dim - 1);
  const int grid = (int)((items + CPB - 1) / CPB);
  mass_kernel<dim, n, T><<<grid, NT, 0, mf->ctx->stream>>>(P, items, (const T *)src->d, (T *)dst->d, add ? 1 : 0);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int mass_by_degree(int degree, dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  switch (degree) {
    case 1: return launch_mass<dim, 2, T>(op, dst, src, add, inverse);
    case 2: return launch_mass<dim, 3, T>(op, dst, src, add, inverse);
    case 3: return launch_mass<dim, 4, T>(op, dst, src, add, inverse);
    case 4: return launch_mass<dim, 5, T>(op, dst, src, add, inverse);
    case 5: return launch_mass<dim, 6, T>(op, dst, src, add, inverse);
    case 6: return launch_mass<dim, 7, T>(op, dst, src, add, inverse);
    case 7: return launch_mass<dim, 8, T>(op, dst, src, add, inverse);
    case 8: return launch_mass<dim, 9, T>(op, dst, src, add, inverse);
    default: set_error("degree must be in 1..8"); return DGX_ERR_UNSUPPORTED;
  }
}
template <int dim, typename T>
int mass_cg_front_by_degree(dgx_op *op, dgx_vec *p, const dgx_vec *r, double beta, bool first, dgx_vec *v, double *pv) {
  switch (op->degree) {
    case 1: return launch_mass_cg_front<dim, 2, T>(op, p, r, beta, first, v, pv);
    case 2: return launch_mass_cg_front<dim, 3, T>(op, p, r, beta, first, v, pv);
    case 3: return launch_mass_cg_front<dim, 4, T>(op, p, r, beta, first, v, pv);
    case 4: return launch_mass_cg_front<dim, 5, T>(op, p, r, beta, first, v, pv);
    case 5: return launch_mass_cg_front<dim, 6, T>(op, p, r, beta, first, v, pv);
    case 6: return launch_mass_cg_front<dim, 7, T>(op, p, r, beta, first, v, pv);
    default: return DGX_ERR_UNSUPPORTED;
  }
}
}  // namespace

// p = r + beta p (p = r when first), v = M p, *pv = p . v in one kernel; DGX_ERR_UNSUPPORTED: caller takes the unfused route
int mass_cg_front(dgx_op *op, dgx_vec *p, const dgx_vec *r, double beta, bool first, dgx_vec *v, double *pv) {
  dgx_mf *mf = op->mf;
  if (op->kind != DGX_OP_MASS || p->dtype != r->dtype || p->dtype != v->dtype || mf->general) return DGX_ERR_UNSUPPORTED;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? mass_cg_front_by_degree<2, double>(op, p, r, beta, first, v, pv) : mass_cg_front_by_degree<2, float>(op, p, r, beta, first, v, pv);
  return mf->dtype == DGX_F64 ? mass_cg_front_by_degree<3, double>(op, p, r, beta, first, v, pv) : mass_cg_front_by_degree<3, float>(op, p, r, beta, first, v, pv);
}

int mass_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add) { return mass_apply_any(op, dst, src, add, false); }

// dst = M^-1 src: the cell mass matrix is (h_x M) (x) (h_y M) (x) (h_z M), so its inverse is the same three sweeps with
// M^-1 / h_d -- the exact solution of project_grad's CG solve (NS.cc:2574-2576) in one kernel.
int mass_inverse_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src) { return mass_apply_any(op, dst, src, false, true); }

int mass_apply_any(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse) {
  dgx_mf *mf = op->mf;
  if (mf->general) return mass_general_apply(op, dst, src, add, inverse);
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? mass_by_degree<2, double>(op->degree, op, dst, src, add, inverse) : mass_by_degree<2, float>(op->degree, op, dst, src, add, inverse);
  return mf->dtype == DGX_F64 ? mass_by_degree<3, double>(op->degree, op, dst, src, add, inverse) : mass_by_degree<3, float>(op->degree, op, dst, src, add, inverse);
}

}  // namespace dgx
