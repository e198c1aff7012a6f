// Pressure operator, NS_stage == 2 branch of apply_add (NS.cc:1407-1414) = assemble_cell_term_pressure
// (NS.cc:1230-1252) + assemble_face_term_pressure (:1259-1292) + assemble_boundary_term_pressure
// (:1299-1326), and its diagonal (compute_diagonal NS.cc:2018-2060 with assemble_diagonal_*_pressure
// :1857-2008), on Cartesian cells.
//
// This file holds the GENERIC kernel: any dim (2, 3), degree 1..8, fp64 / fp32, every boundary type,
// any number of cells, ghost neighbours.  It is cell-centric (each cell gathers the traces of its
// 2*dim neighbours and writes its own dofs once: no atomics, no compress(add), deterministic) and
// uses the Kronecker-factored form documented in pressure_op.cuh.  The blocked fast path for large
// 3-D meshes lives in pressure_fast.cu.
#include <cstdlib>
#include <cstring>

#include "pressure_op.cuh"
#include "reduce.cuh"

namespace dgx {

double pressure_kappa(const dgx_op *op) {
  // coeff exactly as NS.cc:1237 (same association of the products)
  const double gamma = gamma_trbdf2(), dt = op->dt;
  const double coeff = (op->stage == 1) ? 1.0e6 * gamma * dt * gamma * dt : 1.0e6 * (1.0 - gamma) * dt * (1.0 - gamma) * dt;
  return 1.0 / coeff;
}

namespace {

// Shared-memory boxes pad every x-row to RS entries when n is even: with the natural stride the x- and y-lines of a warp
// fall on few banks (n = 4: 4-way conflicts, more than half of all shared-memory wavefronts of the fp32 level operator).
template <int n> __host__ __device__ constexpr int row_stride() { return n % 2 == 0 ? n + 1 : n; }

template <int dim, int n>
__device__ __forceinline__ void line_base_stride(int d, int l, int &base, int &stride) {   // global (unpadded) layout
  if (dim == 3) {
    if (d == 0) { base = l * n; stride = 1; }
    else if (d == 1) { base = (l % n) + n * n * (l / n); stride = n; }
    else { base = l; stride = n * n; }
  } else {
    if (d == 0) { base = l * n; stride = 1; }
    else { base = l; stride = n; }
  }
}
template <int dim, int n>
__device__ __forceinline__ void padded_base_stride(int d, int l, int &base, int &stride) {   // shared-memory box
  constexpr int RS = row_stride<n>();
  if (dim == 3) {
    if (d == 0) { base = l * RS; stride = 1; }
    else if (d == 1) { base = (l % n) + RS * n * (l / n); stride = RS; }
    else { base = (l % n) + RS * (l / n); stride = RS * n; }
  } else {
    if (d == 0) { base = l * RS; stride = 1; }
    else { base = l; stride = RS; }
  }
}

template <int dim, int n>
__host__ __device__ constexpr int cells_per_block() {
  return (192 / ipow(n, dim - 1)) > 0 ? (192 / ipow(n, dim - 1)) : 1;
}

// one batch of CPB cells: dst[cells] (+)= A src.  Shared by the generic kernel (one batch per block) and the
// single-CTA coarse-grid CG kernel (loops over the batches of a tiny level).
template <int dim, int n, typename T>
__device__ __forceinline__ void pressure_cell_batch(const PresParams<T, n> &P, const int *nbr, int n_owned, const T *src, T *dst, int add,
                                                    int cell0, T *su, T *sa, const Epilogue<T> &E = Epilogue<T>(), int cell_end = -1) {
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = cells_per_block<dim, n>(), NT = CPB * NL;
  constexpr int RS = row_stride<n>(), NBX = RS * NL;
  auto pidx = [](int idx) { return idx + (RS - n) * (idx / n); };   // block-linear dof index -> padded shared-memory index
  const int tid = threadIdx.x;
  const int nblk = min(CPB, (cell_end < 0 ? n_owned : cell_end) - cell0);
  for (int idx = tid; idx < nblk * ND; idx += NT) su[pidx(idx)] = src[(size_t)cell0 * ND + idx];
  __syncthreads();
  const int c = tid / NL, l = tid % NL;
  const bool active = c < nblk;
  const int cell = cell0 + c;
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    if (active) {
      int base, stride, pb, ps;
      line_base_stride<dim, n>(d, l, base, stride);
      padded_base_stride<dim, n>(d, l, pb, ps);
      // both neighbour lines first: their latency overlaps the own-line work
      int nb[2];
      T xn[2][n];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        nb[s] = nbr[(size_t)(2 * d + s) * n_owned + cell];
        const T *un = src + (size_t)max(nb[s], 0) * ND + base;   // boundary sides read cell 0 and drop it
        if (d == 0) load_row<T, n>(un, xn[s]);
        else {
#pragma unroll
          for (int j = 0; j < n; ++j) xn[s][j] = un[stride * j];
        }
      }
      T u[n], t[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = su[c * NBX + pb + ps * i];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) acc += P.Kt[d][i * n + j] * u[j];
        t[i] = acc;
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (nb[s] >= 0 || nb[s] == -2) {
          const T V = u[s * (n - 1)];
          T G = 0;
#pragma unroll
          for (int j = 0; j < n; ++j) G += P.g[s][j] * u[j];
          T FV, FG;
          if (nb[s] >= 0) {
            T Gn = 0;
#pragma unroll
            for (int j = 0; j < n; ++j) Gn += P.gn[s][j] * xn[s][j];
            const T Vn = xn[s][(1 - s) * (n - 1)];
            FV = T(-0.5) * (G + Gn) + P.C * (V - Vn);
            FG = T(-0.5) * (V - Vn);
          } else {   // boundary_id == 1: weak Dirichlet (NS.cc:1308-1322)
            FV = -G + P.C * V;
            FG = -V;
          }
#pragma unroll
          for (int i = 0; i < n; ++i) t[i] += P.a[d][s][i] * FV + P.b[d][s][i] * FG;
        }
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const int o = c * NBX + pb + ps * i;
        if (d == 0) sa[o] = P.kappa * u[i] + t[i];
        else sa[o] += t[i];
      }
    }
    __syncthreads();
  }
  // mass: (h_x M) (x) (h_y M) (x) (h_z M)
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    if (active) {
      int pb, ps;
      padded_base_stride<dim, n>(d, l, pb, ps);
      T u[n];
#pragma unroll
      for (int i = 0; i < n; ++i) u[i] = sa[c * NBX + pb + ps * i];
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = 0;
#pragma unroll
        for (int j = 0; j < n; ++j) acc += P.Mh[d][i * n + j] * u[j];
        sa[c * NBX + pb + ps * i] = acc;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nblk * ND; idx += NT) {
    const size_t o = (size_t)cell0 * ND + idx;
    const int pi = pidx(idx);
    if (E.on) {
      T v = E.c * (E.rhs[o] - sa[pi]);
      if (E.d) v *= E.d[o];
      v += E.a * su[pi];
      if (E.xold) v += E.b * E.xold[o];
      dst[o] = v;
    } else {
      dst[o] = add ? dst[o] + sa[pi] : sa[pi];
    }
  }
}

template <int dim, int n, typename T>
__global__ void __launch_bounds__(cells_per_block<dim, n>() * ipow(n, dim - 1))
pressure_generic_kernel(const __grid_constant__ PresParams<T, n> P, const int *__restrict__ nbr, int n_owned,
                        const T *__restrict__ src, T *__restrict__ dst, int add, const __grid_constant__ Epilogue<T> E, int cell_begin,
                        int cell_end) {
  constexpr int NBX = row_stride<n>() * ipow(n, dim - 1), CPB = cells_per_block<dim, n>();
  __shared__ T su[CPB * NBX];
  __shared__ T sa[CPB * NBX];
  pressure_cell_batch<dim, n, T>(P, nbr, n_owned, src, dst, add, cell_begin + blockIdx.x * CPB, su, sa, E, cell_end);
}

// Q1 (n = 2): a cell has 2^dim dofs (64 B in 3-D fp64), so the line-per-thread decomposition of the generic kernel leaves 4 threads per
// cell shuffling 8 values through shared memory.  Here ONE THREAD owns a cell: its dofs and the accumulator live in registers, same
// algebra and tables as pressure_cell_batch, no barrier but one.  A warp reading "my cell" with per-thread 16-byte loads touches 16
// cache lines per instruction (64 B stride between lanes: the first version was bound by L1 tag look-ups, 75 % l1tex with 5 % LSU
// issue), so the block's 128 consecutive cells are fetched with lane-contiguous 16-byte loads into a padded shared-memory copy, from
// which every thread takes its own cell AND those neighbour cells that lie inside the block (most of them: consecutive cells of the
// hierarchical order form a compact brick); only neighbours outside the block are read from global memory.
template <int dim, typename T>
__global__ void __launch_bounds__(128, 6)
pressure_q1_kernel(const __grid_constant__ PresParams<T, 2> P, const int *__restrict__ nbr, int n_owned, const T *__restrict__ src, T *__restrict__ dst,
                   int add, const __grid_constant__ Epilogue<T> E, int cell_begin, int cell_end) {
  constexpr int ND = 1 << dim, VPC = ND * (int)sizeof(T) / 16, RW = VPC + 1;   // 16-byte vectors per cell; row stride (one vector of padding)
  __shared__ int4 sm4[128 * RW];
  const int bcell0 = cell_begin + blockIdx.x * 128, ncb = min(128, cell_end - bcell0);
  {
    const int4 *g = reinterpret_cast<const int4 *>(src + (size_t)bcell0 * ND);
    for (int v = threadIdx.x; v < ncb * VPC; v += 128) sm4[(v / VPC) * RW + (v % VPC)] = __ldg(g + v);
  }
  __syncthreads();
  const int cell = bcell0 + threadIdx.x;
  if (cell >= cell_end) return;
  auto cell_from_smem = [&](int row, T (&out)[ND]) {
#pragma unroll
    for (int k = 0; k < VPC; ++k) {
      const int4 t = sm4[row * RW + k];
      T tmp[16 / sizeof(T)];
      memcpy(tmp, &t, 16);
#pragma unroll
      for (int j = 0; j < (int)(16 / sizeof(T)); ++j) out[k * (16 / sizeof(T)) + j] = tmp[j];
    }
  };
  T u[ND], acc[ND];
  cell_from_smem(threadIdx.x, u);
#pragma unroll
  for (int i = 0; i < ND; ++i) acc[i] = P.kappa * u[i];
  // neighbour indices of all faces first (independent loads), then one neighbour cell at a time: 8 values of neighbour data live
  int nbs[2 * dim];
#pragma unroll
  for (int f = 0; f < 2 * dim; ++f) nbs[f] = nbr[(size_t)f * n_owned + cell];
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    const int st = 1 << d;
#pragma unroll
    for (int i0 = 0; i0 < ND; ++i0) {
      if (i0 & st) continue;             // lines along d: (i0, i0 + st) with bit d of i0 clear
      const int i1 = i0 + st;
      acc[i0] += P.Kt[d][0] * u[i0] + P.Kt[d][1] * u[i1];
      acc[i1] += P.Kt[d][2] * u[i0] + P.Kt[d][3] * u[i1];
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int nb = nbs[2 * d + s];
      if (!(nb >= 0 || nb == -2)) continue;
      T xn[ND];
      if (nb >= 0) {
        const int row = nb - bcell0;
        if (row >= 0 && row < ncb) cell_from_smem(row, xn);
        else load_row<T, ND>(src + (size_t)nb * ND, xn);
      }
#pragma unroll
      for (int i0 = 0; i0 < ND; ++i0) {
        if (i0 & st) continue;
        const int i1 = i0 + st;
        const T u0 = u[i0], u1 = u[i1];
        const T V = s ? u1 : u0;
        const T G = P.g[s][0] * u0 + P.g[s][1] * u1;
        T FV, FG;
        if (nb >= 0) {
          const T x0 = xn[i0], x1 = xn[i1];
          const T Gn = P.gn[s][0] * x0 + P.gn[s][1] * x1;
          const T Vn = s ? x0 : x1;    // the neighbour's value on the shared face
          FV = T(-0.5) * (G + Gn) + P.C * (V - Vn);
          FG = T(-0.5) * (V - Vn);
        } else {                       // boundary_id == 1: weak Dirichlet (NS.cc:1308-1322)
          FV = -G + P.C * V;
          FG = -V;
        }
        acc[i0] += P.a[d][s][0] * FV + P.b[d][s][0] * FG;
        acc[i1] += P.a[d][s][1] * FV + P.b[d][s][1] * FG;
      }
    }
  }
  // mass: (h_x M) (x) (h_y M) (x) (h_z M)
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    const int st = 1 << d;
#pragma unroll
    for (int i0 = 0; i0 < ND; ++i0) {
      if (i0 & st) continue;
      const int i1 = i0 + st;
      const T a0 = acc[i0], a1 = acc[i1];
      acc[i0] = P.Mh[d][0] * a0 + P.Mh[d][1] * a1;
      acc[i1] = P.Mh[d][2] * a0 + P.Mh[d][3] * a1;
    }
  }
  const size_t o = (size_t)cell * ND;
  if (E.on) {
    T rhs[ND], dd[ND], xo[ND];
    load_row<T, ND>(E.rhs + o, rhs);
    if (E.d) load_row<T, ND>(E.d + o, dd);
    if (E.xold) load_row<T, ND>(E.xold + o, xo);
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      T v = E.c * (rhs[i] - acc[i]);
      if (E.d) v *= dd[i];
      v += E.a * u[i];
      if (E.xold) v += E.b * xo[i];
      acc[i] = v;
    }
    store_row<T, ND>(dst + o, acc, 0);
  } else {
    store_row<T, ND>(dst + o, acc, add);
  }
}

// MGCoarseGridIterativeSolver<SolverCG, PreconditionIdentity> (NS.cc:2519-2529) for a level of a few cells, entirely inside
// ONE CTA: no launch and no host round trip per iteration.  Same algebra as cg_core (solvers.cu) with x0 = 0.
// info[0] = iterations, info[1] = 1 if converged; res[0] = final residual norm.
template <int dim, int n, typename T>
__global__ void __launch_bounds__(cells_per_block<dim, n>() * ipow(n, dim - 1))
pressure_coarse_cg_kernel(const __grid_constant__ PresParams<T, n> P, const int *nbr, int n_cells, const T *b, T *x, T *r, T *p, T *v,
                          const double *tol_ptr, int max_it, int *info, double *res_out, int *log_count, int log_cap) {
  // tolerance and log slot come from device memory: the launch is identical from solve to solve (CUDA-graph replay)
  const double tol = *tol_ptr;
  constexpr int NL = ipow(n, dim - 1), ND = NL * n, CPB = cells_per_block<dim, n>(), NT = CPB * NL;
  __shared__ T su[CPB * row_stride<n>() * NL];
  __shared__ T sa[CPB * row_stride<n>() * NL];
  __shared__ double red[32];
  __shared__ double bc;
  const int tid = threadIdx.x, N = n_cells * ND;
  auto block_sum = [&](double val) -> double {
    val = warp_sum_existing(val);   // NT is not a multiple of 32 for most degrees
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    if (tid == 0) {
      double s = 0;
      for (int w = 0; w < (NT + 31) / 32; ++w) s += red[w];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  double acc = 0;
  for (int i = tid; i < N; i += NT) { const T bi = b[i]; r[i] = bi; x[i] = 0; acc += (double)bi * (double)bi; }
  double rr = block_sum(acc), res = sqrt(rr), prev_rr = 0;
  int it = 0;
  bool conv = res <= tol;
  while (!conv && it < max_it && isfinite(res)) {
    ++it;
    if (it > 1) {
      const T beta = (T)(rr / prev_rr);
      for (int i = tid; i < N; i += NT) p[i] = beta * p[i] + r[i];
    } else {
      for (int i = tid; i < N; i += NT) p[i] = r[i];
    }
    __syncthreads();
    for (int c0 = 0; c0 < n_cells; c0 += CPB) {
      pressure_cell_batch<dim, n, T>(P, nbr, n_cells, p, v, 0, c0, su, sa);
      __syncthreads();
    }
    acc = 0;
    for (int i = tid; i < N; i += NT) acc += (double)p[i] * (double)v[i];
    const double pv = block_sum(acc);
    const T alpha = (T)(rr / pv);
    acc = 0;
    for (int i = tid; i < N; i += NT) {
      x[i] += alpha * p[i];
      const T ri = r[i] - alpha * v[i];
      r[i] = ri;
      acc += (double)ri * (double)ri;
    }
    prev_rr = rr;
    rr = block_sum(acc);
    res = sqrt(rr);
    conv = res <= tol;
  }
  if (tid == 0) {   // append to the device-side log: (iterations, converged), residual
    const int slot = atomicAdd(log_count, 1) % log_cap;
    info[2 * slot] = it; info[2 * slot + 1] = conv ? 1 : 0; res_out[slot] = res;
  }
}

// Inverse "diagonal" exactly as the reference probes it (NS.cc:1857-2008): the interior-face probe sets
// local dof i on BOTH sides (NS.cc:1921-1922), so the neighbour's i-th basis function leaks into the
// entry; the boundary probe uses 2*sigma (NS.cc:1997) although the operator uses sigma (NS.cc:1320).
template <typename T, int n>
struct DiagParams {
  T Kd[n];          // K_ii
  T Md[n];          // M_ii
  T qi[2][n];       // interior-face probe, side s
  T qd[2][n];       // id == 1 boundary probe, side s
  T ih2[MAX_DIM];   // 1/h_d^2
  T vol, kappa;
};

template <int dim, int n, typename T>
__global__ void pressure_diag_kernel(const __grid_constant__ DiagParams<T, n> P, const int *__restrict__ nbr, int n_owned, T *__restrict__ out) {
  constexpr int ND = ipow(n, dim);
  const long long total = (long long)n_owned * ND;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(g / ND);
    int r = (int)(g - (long long)cell * ND);
    int idx[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < dim; ++d) { idx[d] = r % n; r /= n; }
    T mprod = 1;
#pragma unroll
    for (int d = 0; d < dim; ++d) mprod *= P.Md[idx[d]];
    T sum = P.kappa * mprod;
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      T l = P.Kd[idx[d]];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int nb = nbr[(size_t)(2 * d + s) * n_owned + cell];
        if (nb >= 0) l += P.qi[s][idx[d]];
        else if (nb == -2) l += P.qd[s][idx[d]];
      }
      T others = 1;
#pragma unroll
      for (int e = 0; e < dim; ++e) if (e != d) others *= P.Md[idx[e]];
      sum += P.ih2[d] * l * others;
    }
    out[g] = T(1) / (P.vol * sum);
  }
}

struct CellRange { long long begin = 0, end = -1; cudaStream_t stream = nullptr; };

template <int dim, int n, typename T>
int launch_generic(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu = nullptr, const CellRange *range = nullptr) {
  dgx_mf *mf = op->mf;
  PresParams<T, n> P;
  fill_pres_params<T, n>(P, op, pressure_kappa(op));
  constexpr int CPB = cells_per_block<dim, n>(), NT = CPB * ipow(n, dim - 1);
  const int n_owned = (int)mf->n_owned;
  if (n_owned == 0) return DGX_OK;
  const int cb = range ? (int)range->begin : 0, ce = range ? (int)range->end : n_owned;
  if (ce <= cb) return DGX_OK;
  const int grid = (ce - cb + CPB - 1) / CPB;
  cudaStream_t stream = range && range->stream ? range->stream : mf->ctx->stream;
  Epilogue<T> E;
  if (fu) {
    E.on = 1;
    E.xold = fu->xold ? (const T *)fu->xold->d : nullptr;
    E.d = fu->d ? (const T *)fu->d->d : nullptr;
    E.rhs = (const T *)fu->rhs->d;
    E.a = (T)fu->a; E.b = (T)fu->b; E.c = (T)fu->c;
  }
  static const bool q1_off = getenv("DGX_NO_Q1_KERNEL") != nullptr;
  if constexpr (n == 2) {
    if (!q1_off) {
      pressure_q1_kernel<dim, T><<<(ce - cb + 127) / 128, 128, 0, stream>>>(P, mf->d_nbr, n_owned, (const T *)src->d, (T *)dst->d, add ? 1 : 0, E, cb, ce);
      count_launch();
      DGX_CUDA_CHECK(cudaGetLastError());
      return DGX_OK;
    }
  }
  pressure_generic_kernel<dim, n, T><<<grid, NT, 0, stream>>>(P, mf->d_nbr, n_owned, (const T *)src->d, (T *)dst->d, add ? 1 : 0, E, cb, ce);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, int n, typename T>
int launch_diag(dgx_op *op, dgx_vec *out) {
  dgx_mf *mf = op->mf;
  const Shape1D &s = get_shape(n - 1, n);
  DiagParams<T, n> P;
  const double C = (double)n * n;
  for (int i = 0; i < n; ++i) {
    P.Kd[i] = (T)s.K[i * n + i];
    P.Md[i] = (T)s.M[i * n + i];
    for (int sd = 0; sd < 2; ++sd) {
      const double ns = sd ? 1.0 : -1.0;
      const double f = s.f[sd][i], g = s.g[sd][i], fo = s.f[1 - sd][i], go = s.g[1 - sd][i];
      // probe u = phi_i on this cell AND phi_i on the neighbour (same local index)
      const double FV = -0.5 * ns * (g + go) + C * (f - fo), FG = -0.5 * (f - fo);
      P.qi[sd][i] = (T)(f * FV + ns * g * FG);
      const double FVd = -ns * g + 2.0 * C * f, FGd = -f;
      P.qd[sd][i] = (T)(f * FVd + ns * g * FGd);
    }
  }
  double vol = 1;
  for (int d = 0; d < MAX_DIM; ++d) {
    const double h = d < dim ? mf->h[d] : 1.0;
    P.ih2[d] = (T)(1.0 / (h * h));
    if (d < dim) vol *= h;
  }
  P.vol = (T)vol;
  P.kappa = (T)pressure_kappa(op);
  const long long total = mf->n_owned * ipow(n, dim);
  if (total == 0) return DGX_OK;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pressure_diag_kernel<dim, n, T><<<grid, 256, 0, mf->ctx->stream>>>(P, mf->d_nbr, (int)mf->n_owned, (T *)out->d);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

#define DGX_DEGREE_SWITCH(CALL)              \
  switch (degree) {                          \
    case 1: return CALL(2);                  \
    case 2: return CALL(3);                  \
    case 3: return CALL(4);                  \
    case 4: return CALL(5);                  \
    case 5: return CALL(6);                  \
    case 6: return CALL(7);                  \
    case 7: return CALL(8);                  \
    case 8: return CALL(9);                  \
    default: set_error("degree must be in 1..8"); return DGX_ERR_UNSUPPORTED; \
  }

template <int dim, typename T>
int generic_by_degree(int degree, dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu = nullptr, const CellRange *range = nullptr) {
#define CALL(N) launch_generic<dim, N, T>(op, dst, src, add, fu, range)
  DGX_DEGREE_SWITCH(CALL)
#undef CALL
}

template <int dim, typename T>
int diag_by_degree(int degree, dgx_op *op, dgx_vec *out) {
#define CALL(N) launch_diag<dim, N, T>(op, out)
  DGX_DEGREE_SWITCH(CALL)
#undef CALL
}

}  // namespace

template <int dim, int n, typename T>
int launch_coarse_cg(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res, int *d_count, int log_cap) {
  dgx_mf *mf = op->mf;
  PresParams<T, n> P;
  fill_pres_params<T, n>(P, op, pressure_kappa(op));
  constexpr int CPB = cells_per_block<dim, n>(), NT = CPB * ipow(n, dim - 1);
  pressure_coarse_cg_kernel<dim, n, T><<<1, NT, 0, mf->ctx->stream>>>(P, mf->d_nbr, (int)mf->n_owned, (const T *)b->d, (T *)x->d, (T *)r->d, (T *)p->d,
                                                                      (T *)v->d, d_tol, max_it, d_info, d_res, d_count, log_cap);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int coarse_by_degree(int degree, dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res, int *d_count, int log_cap) {
#define CALL(N) launch_coarse_cg<dim, N, T>(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap)
  DGX_DEGREE_SWITCH(CALL)
#undef CALL
}

// single-CTA coarse solve; DGX_ERR_UNSUPPORTED when the level is not tiny / neither on one rank nor replicated on all ranks
// (caller falls back to cg_core)
int pressure_coarse_cg(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res, int *d_count, int log_cap) {
  dgx_mf *mf = op->mf;
  if (op->kind != DGX_OP_PRESSURE || (mf->ctx->nranks != 1 && !mf->replicated) || mf->n_ghost != 0 || mf->n_owned < 1 || mf->n_owned > 64) return DGX_ERR_UNSUPPORTED;
  if (mf->general) return pressure_general_coarse_cg(op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
  const int degree = op->degree;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? coarse_by_degree<2, double>(degree, op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap)
                                                 : coarse_by_degree<2, float>(degree, op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
  return mf->dtype == DGX_F64 ? coarse_by_degree<3, double>(degree, op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap)
                              : coarse_by_degree<3, float>(degree, op, x, b, r, p, v, d_tol, max_it, d_info, d_res, d_count, log_cap);
}

int pressure_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add) {
  dgx_mf *mf = op->mf;
  if (mf->general) return pressure_general_apply(op, dst, src, add, nullptr, 1.0);
  if (op->variant != 1) {
    int rc = pressure_apply_fast(op, dst, src, add);
    if (rc != DGX_ERR_UNSUPPORTED) return rc;
    if (op->variant == 2) return rc;
  }
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? generic_by_degree<2, double>(op->degree, op, dst, src, add)
                                                 : generic_by_degree<2, float>(op->degree, op, dst, src, add);
  return mf->dtype == DGX_F64 ? generic_by_degree<3, double>(op->degree, op, dst, src, add)
                              : generic_by_degree<3, float>(op->degree, op, dst, src, add);
}

int pressure_apply_range(dgx_op *op, dgx_vec *dst, const dgx_vec *src, long long cell_begin, long long cell_end, cudaStream_t stream) {
  dgx_mf *mf = op->mf;
  if (op->kind != DGX_OP_PRESSURE || mf->general) return DGX_ERR_UNSUPPORTED;   // several ranks: the caller has imported the ghosts
  CellRange r;
  r.begin = cell_begin; r.end = cell_end; r.stream = stream;
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? generic_by_degree<2, double>(op->degree, op, dst, src, false, nullptr, &r)
                                                 : generic_by_degree<2, float>(op->degree, op, dst, src, false, nullptr, &r);
  return mf->dtype == DGX_F64 ? generic_by_degree<3, double>(op->degree, op, dst, src, false, nullptr, &r)
                              : generic_by_degree<3, float>(op->degree, op, dst, src, false, nullptr, &r);
}

// dst = a x + b xold + c d .* (rhs - A x) in ONE kernel (generic path).  dst must differ from x (neighbour cells read x).
int pressure_apply_fused(dgx_op *op, dgx_vec *dst, dgx_vec *x, const FusedUpdate &fu) {
  dgx_mf *mf = op->mf;
  if (op->kind != DGX_OP_PRESSURE || dst->d == x->d) return DGX_ERR_UNSUPPORTED;
  int rc = ghost_exchange(x);
  if (rc) return rc;
  if (mf->general) return pressure_general_apply(op, dst, x, false, &fu, 1.0);
  // The marching kernel applies the same epilogue in its Z warps (pressure_apply_fast_fused).  For even n the update vectors
  // reach them through bulk copies one plane ahead; for odd n they would be plain loads on the producer warps' critical
  // path (2.1 ms per fused step against 1.0 ms here, 128^3 Q3 fp32 before the staging): there it stays opt-in (variant 2).
  if (op->variant == 2 || (op->variant == 0 && op->degree == 3 && !getenv("DGX_NO_FUSED_MARCH"))) {
    rc = pressure_apply_fast_fused(op, dst, x, fu);
    if (rc != DGX_ERR_UNSUPPORTED) return rc;
  }
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? generic_by_degree<2, double>(op->degree, op, dst, x, false, &fu)
                                                 : generic_by_degree<2, float>(op->degree, op, dst, x, false, &fu);
  return mf->dtype == DGX_F64 ? generic_by_degree<3, double>(op->degree, op, dst, x, false, &fu)
                              : generic_by_degree<3, float>(op->degree, op, dst, x, false, &fu);
}

int pressure_diagonal(dgx_op *op, dgx_vec *out) {
  dgx_mf *mf = op->mf;
  if (mf->general) return pressure_general_diagonal(op, out);
  if (mf->dim == 2) return mf->dtype == DGX_F64 ? diag_by_degree<2, double>(op->degree, op, out) : diag_by_degree<2, float>(op->degree, op, out);
  return mf->dtype == DGX_F64 ? diag_by_degree<3, double>(op->degree, op, out) : diag_by_degree<3, float>(op->degree, op, out);
}

}  // namespace dgx
