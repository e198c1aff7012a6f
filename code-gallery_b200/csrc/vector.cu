// Device vector: the subset of LinearAlgebra::distributed::Vector<Number> the hot path uses
// (call sites NS.cc:2408-2415 equ, :2410/2954/2962 +=/-=, :2450/2486/2574 l2_norm,
// :2585/2597 linfty_norm, :452/790/917 update_ghost_values; plus what SolverCG / SolverGMRES /
// PreconditionChebyshev call inside deal.II: operator*, add, sadd, add_and_dot, mean_value).
// All kernels are HBM-bound streaming passes: 16-byte vector loads, grid = multiple of the SM count.
// Reductions accumulate in double, per-block partials are combined in a fixed order by the last
// block (deterministic), then all-reduced over NCCL.
#include <cstring>

#include "dgx_internal.h"

namespace dgx {
int comm_exchange(dgx_ctx *ctx, cudaStream_t stream, int npeers, const int *ranks, const void *const *sendbuf, const size_t *sendbytes,
                  void *const *recvbuf, const size_t *recvbytes);

namespace {
constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

inline int grid_for(long long n, int per_thread = 4) {
  long long b = (n + (long long)kThreads * per_thread - 1) / ((long long)kThreads * per_thread);
  if (b < 1) b = 1;
  if (b > kMaxBlocks) b = kMaxBlocks;
  return (int)b;
}

template <typename T, typename F>
__global__ void __launch_bounds__(kThreads) map1_kernel(T *__restrict__ dst, long long n, F f) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = f(dst[i], i);
}

template <typename T, typename U, typename F>
__global__ void __launch_bounds__(kThreads) map2_kernel(T *__restrict__ dst, const U *__restrict__ src, long long n, F f) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = f(dst[i], src[i]);
}


// op: 0 = sum a*b, 1 = max |a|, 2 = sum a ; optionally fused update dst += alpha*v before dot(dst, w)
// (upd2 += alpha2 * v2 rides along in the same pass: the x and r updates of one CG iteration)
template <typename T, int OP, bool FUSED_ADD>
__global__ void __launch_bounds__(kThreads) reduce_kernel(const T *a, const T *b, T *upd, const T *v, T alpha, long long n,
                                                          double *partial, double *result, T *upd2 = nullptr, const T *v2 = nullptr,
                                                          T alpha2 = T(0)) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (FUSED_ADD && upd2) upd2[i] += alpha2 * v2[i];
    if (FUSED_ADD) {
      T x = upd[i] + alpha * v[i];
      const T bv = (b == (const T *)upd) ? x : b[i];   // add_and_dot(a, v, *this) aliases
      upd[i] = x;
      acc += (double)x * (double)bv;
    } else if (OP == 0) acc += (double)a[i] * (double)b[i];
    else if (OP == 1) acc = fmax(acc, fabs((double)a[i]));
    else acc += (double)a[i];
  }
  __shared__ double sm[kThreads / 32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) {
    double other = __shfl_down_sync(0xffffffffu, acc, o);
    acc = (OP == 1) ? fmax(acc, other) : acc + other;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = sm[0];
    for (int w = 1; w < kThreads / 32; ++w) s = (OP == 1) ? fmax(s, sm[w]) : s + sm[w];
    partial[blockIdx.x] = s;
    __threadfence();
    unsigned int t = atomicInc(scratch_ticket(partial), gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    // fixed-order tree over the partials
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {
      double p = ((volatile double *)partial)[i];
      s = (OP == 1) ? fmax(s, p) : s + p;
    }
    for (int o = 16; o > 0; o >>= 1) {
      double other = __shfl_down_sync(0xffffffffu, s, o);
      s = (OP == 1) ? fmax(s, other) : s + other;
    }
    if (threadIdx.x == 0) *result = s;
  }
}

template <typename T>
__global__ void pack_cells_kernel(T *__restrict__ out, const T *__restrict__ v, const int *__restrict__ cells, long long ncells, int dpc) {
  const long long total = ncells * dpc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long c = i / dpc;
    int k = (int)(i - c * dpc);
    out[i] = v[(long long)cells[c] * dpc + k];
  }
}

// inverse of pack_cells_kernel: v[cells[c]] = in[c] (whole cells)
template <typename T>
__global__ void scatter_cells_kernel(T *__restrict__ v, const T *__restrict__ in, const int *__restrict__ cells, long long ncells, int dpc) {
  const long long total = ncells * dpc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long c = i / dpc;
    int k = (int)(i - c * dpc);
    v[(long long)cells[c] * dpc + k] = in[i];
  }
}

constexpr int kMaxMulti = 32;
template <typename T> struct PtrTable { const T *v[kMaxMulti]; double h[kMaxMulti]; };

// out[j] = sum_i w[i] * V_j[i], j < k: ONE pass over w for all k dot products of a Gram-Schmidt sweep (SolverGMRES).
// Two elements per thread and iteration (16-byte loads for double), grid = 8 blocks per SM: with one element per
// iteration and 2 blocks per SM the pass ran at 1.5 TB/s, bound by the loads in flight.
template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

template <typename T, int KB>
__global__ void __launch_bounds__(kThreads) multi_dot_kernel(const T *__restrict__ w, const __grid_constant__ PtrTable<T> tab, int k0, int k, long long n,
                                                             double *partial, double *result) {
  using P2 = typename Pair<T>::type;
  double acc[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) acc[j] = 0.0;
  const int kk = min(KB, k - k0);
  const long long n2 = n / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const P2 wi = reinterpret_cast<const P2 *>(w)[i];
#pragma unroll
    for (int j = 0; j < KB; ++j)
      if (j < kk) {
        const P2 vj = reinterpret_cast<const P2 *>(tab.v[k0 + j])[i];
        acc[j] += (double)wi.x * (double)vj.x + (double)wi.y * (double)vj.y;
      }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    for (int j = 0; j < kk; ++j) acc[j] += (double)w[n - 1] * (double)tab.v[k0 + j][n - 1];
  }
  __shared__ double sm[KB][kThreads / 32];
  __shared__ bool last;
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    double a = acc[j];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sm[j][threadIdx.x >> 5] = a;
  }
  __syncthreads();
  if (threadIdx.x < KB) {
    double s = 0;
    for (int wv = 0; wv < kThreads / 32; ++wv) s += sm[threadIdx.x][wv];
    partial[(size_t)blockIdx.x * KB + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicInc(scratch_ticket(partial), gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (last) {   // fixed-order combination of the block partials: warp j sums column j
    __threadfence();
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (j < KB) {
      double s = 0;
      for (int b = lane; b < (int)gridDim.x; b += 32) s += ((volatile double *)partial)[(size_t)b * KB + j];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0 && k0 + j < k) result[k0 + j] = s;
    }
  }
}

// w -= sum_j h_j V_j in one pass
template <typename T>
__global__ void __launch_bounds__(kThreads) multi_axpy_kernel(T *__restrict__ w, const __grid_constant__ PtrTable<T> tab, int k, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double x = (double)w[i];
    for (int j = 0; j < k; ++j) x -= tab.h[j] * (double)tab.v[j][i];
    w[i] = (T)x;
  }
}

template <int OP, bool FUSED>
int reduce_dispatch(const dgx_mf *mf, int dtype, const void *a, const void *b, void *upd, const void *v, double alpha, long long n,
                    double *out, bool is_max, void *upd2 = nullptr, const void *v2 = nullptr, double alpha2 = 0.0) {
  dgx_ctx *ctx = mf->ctx;
  const int grid = grid_for(n, 8);
  double *partial = ctx->d_scratch + kScratchPartial, *result = ctx->d_scratch;
  if (dtype == DGX_F64)
    reduce_kernel<double, OP, FUSED><<<grid, kThreads, 0, ctx->stream>>>((const double *)a, (const double *)b, (double *)upd,
                                                                        (const double *)v, alpha, n, partial, result, (double *)upd2,
                                                                        (const double *)v2, alpha2);
  else
    reduce_kernel<float, OP, FUSED><<<grid, kThreads, 0, ctx->stream>>>((const float *)a, (const float *)b, (float *)upd,
                                                                       (const float *)v, (float)alpha, n, partial, result, (float *)upd2,
                                                                       (const float *)v2, (float)alpha2);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  if (!mf->replicated) {   // a replicated level holds the whole vector on every rank
    int rc = is_max ? allreduce_max(ctx, result, 1) : allreduce_sum(ctx, result, 1);
    if (rc) return rc;
  }
  DGX_CUDA_CHECK(cudaMemcpyAsync(ctx->h_scratch, result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  *out = ctx->h_scratch[0];
  return DGX_OK;
}

bool same_shape(const dgx_vec *a, const dgx_vec *b) { return a && b && a->n_owned == b->n_owned; }
}  // namespace

static int ghost_exchange_on(dgx_vec *v, cudaStream_t stream) {
  dgx_mf *mf = v->mf;
  dgx_ctx *ctx = mf->ctx;
  if (ctx->nranks == 1 || mf->peers.empty()) return DGX_OK;
  {
    int done = 0;
    int rc = p2p_ghost_exchange(v, stream, &done);
    if (rc) return rc;
    if (done) return DGX_OK;
  }
  const int dpc = (int)v->dofs_per_cell;
  const size_t es = v->elem();
  size_t total_send = 0;
  for (auto &p : mf->peers) total_send += p.send_cells.size();
  const size_t need = total_send * dpc * es;
  if (need > v->sendbuf_bytes) {
    if (v->d_sendbuf) cudaFree(v->d_sendbuf);
    DGX_CUDA_CHECK(cudaMalloc(&v->d_sendbuf, need));
    v->sendbuf_bytes = need;
  }
  std::vector<int> ranks;
  std::vector<const void *> sb;
  std::vector<void *> rb;
  std::vector<size_t> sbytes, rbytes;
  size_t off = 0;
  for (auto &p : mf->peers) {
    const long long nc = (long long)p.send_cells.size();
    char *dstp = (char *)v->d_sendbuf + off;
    if (nc) {
      const int grid = grid_for(nc * dpc);
      if (v->dtype == DGX_F64)
        pack_cells_kernel<double><<<grid, kThreads, 0, stream>>>((double *)dstp, (const double *)v->d, p.d_send_cells, nc, dpc);
      else
        pack_cells_kernel<float><<<grid, kThreads, 0, stream>>>((float *)dstp, (const float *)v->d, p.d_send_cells, nc, dpc);
      count_launch();
    }
    ranks.push_back(p.rank);
    sb.push_back(dstp);
    sbytes.push_back((size_t)nc * dpc * es);
    rb.push_back((char *)v->d + ((size_t)v->n_owned + (size_t)p.recv_first * dpc) * es);
    rbytes.push_back((size_t)p.recv_count * dpc * es);
    off += (size_t)nc * dpc * es;
  }
  DGX_CUDA_CHECK(cudaGetLastError());
  return comm_exchange(ctx, stream, (int)ranks.size(), ranks.data(), sb.data(), sbytes.data(), rb.data(), rbytes.data());
}


int ghost_exchange(dgx_vec *v) { return ghost_exchange_on(v, v->mf->ctx->stream); }

// v[cells[c]] = staged[c] for whole cells, on `stream` (the multi-rank host entry point uploads the rank-boundary cells first)
int scatter_cells(dgx_vec *v, const void *d_staged, const int *d_cells, long long ncells, cudaStream_t stream) {
  if (ncells == 0) return DGX_OK;
  const int dpc = (int)v->dofs_per_cell;
  const int grid = grid_for(ncells * dpc);
  if (v->dtype == DGX_F64) scatter_cells_kernel<double><<<grid, kThreads, 0, stream>>>((double *)v->d, (const double *)d_staged, d_cells, ncells, dpc);
  else scatter_cells_kernel<float><<<grid, kThreads, 0, stream>>>((float *)v->d, (const float *)d_staged, d_cells, ncells, dpc);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

int ghost_exchange_start(dgx_vec *v) {
  dgx_ctx *ctx = v->mf->ctx;
  if (ctx->nranks == 1 || v->mf->peers.empty()) return DGX_OK;
  DGX_CUDA_CHECK(cudaEventRecord(ctx->ev_src_ready, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_src_ready, 0));
  int rc = ghost_exchange_on(v, ctx->comm_stream);
  if (rc) return rc;
  DGX_CUDA_CHECK(cudaEventRecord(ctx->ev_ghosts_done, ctx->comm_stream));
  return DGX_OK;
}

int ghost_exchange_wait(dgx_ctx *ctx) {
  if (ctx->nranks == 1) return DGX_OK;
  DGX_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_ghosts_done, 0));
  return DGX_OK;
}

}  // namespace dgx

using namespace dgx;

extern "C" {

int dgx_vec_create(dgx_mf *mf, int degree, int ncomp, dgx_vec **out) {
  DGX_REQUIRE(mf && out, "null argument");
  DGX_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
  DGX_REQUIRE(ncomp >= 1 && ncomp <= 3, "ncomp");
  auto *v = new dgx_vec();
  v->mf = mf; v->degree = degree; v->ncomp = ncomp; v->dtype = mf->dtype;
  long long dpc = ncomp;
  for (int d = 0; d < mf->dim; ++d) dpc *= (degree + 1);
  v->dofs_per_cell = dpc;
  v->n_owned = mf->n_owned * dpc;
  v->n_ghost = mf->n_ghost * dpc;
  v->n_global = mf->n_global * dpc;
  const size_t bytes = (size_t)(v->n_owned + v->n_ghost) * v->elem();
  if (cudaSetDevice(mf->ctx->device) != cudaSuccess || cudaMalloc(&v->d, bytes ? bytes : 8) != cudaSuccess) {
    set_error(std::string("cudaMalloc of vector failed: ") + cudaGetErrorString(cudaGetLastError()));
    delete v;
    return DGX_ERR_CUDA;
  }
  cudaMemsetAsync(v->d, 0, bytes, mf->ctx->stream);
  *out = v;
  return DGX_OK;
}

int dgx_vec_destroy(dgx_vec *v) {
  if (!v || v->borrowed) return DGX_OK;
  if (v->d) cudaFree(v->d);
  if (v->d_sendbuf) cudaFree(v->d_sendbuf);
  delete v;
  return DGX_OK;
}

long long dgx_vec_size(const dgx_vec *v) { return v ? v->n_owned : 0; }
long long dgx_vec_ghost_size(const dgx_vec *v) { return v ? v->n_ghost : 0; }
long long dgx_vec_global_size(const dgx_vec *v) { return v ? v->n_global : 0; }
int dgx_vec_dtype(const dgx_vec *v) { return v ? v->dtype : -1; }
void *dgx_vec_data(dgx_vec *v) { return v ? v->d : nullptr; }

int dgx_vec_copy_from_host(dgx_vec *v, const void *host, long long n) {
  DGX_REQUIRE(v && host && n == v->n_owned, "size mismatch");
  DGX_CUDA_CHECK(cudaMemcpyAsync(v->d, host, (size_t)n * v->elem(), cudaMemcpyHostToDevice, v->mf->ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(v->mf->ctx->stream));
  return DGX_OK;
}

int dgx_vec_copy_to_host(const dgx_vec *v, void *host, long long n) {
  DGX_REQUIRE(v && host && n == v->n_owned, "size mismatch");
  DGX_CUDA_CHECK(cudaMemcpyAsync(host, v->d, (size_t)n * v->elem(), cudaMemcpyDeviceToHost, v->mf->ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(v->mf->ctx->stream));
  return DGX_OK;
}

#define LAUNCH_MAP1(vec, n, ...)                                                                             \
  do {                                                                                                       \
    cudaStream_t st__ = (vec)->mf->ctx->stream;                                                              \
    const int grid__ = grid_for(n);                                                                          \
    if ((vec)->dtype == DGX_F64) { typedef double T; map1_kernel<T><<<grid__, kThreads, 0, st__>>>((T *)(vec)->d, n, __VA_ARGS__); } \
    else { typedef float T; map1_kernel<T><<<grid__, kThreads, 0, st__>>>((T *)(vec)->d, n, __VA_ARGS__); }  \
    count_launch();                                                                                          \
    DGX_CUDA_CHECK(cudaGetLastError());                                                                      \
  } while (0)

int dgx_vec_set(dgx_vec *v, double value) {
  DGX_REQUIRE(v, "null argument");
  const long long n = v->n_owned;
  if (v->dtype == DGX_F64) { double a = value; map1_kernel<double><<<grid_for(n), kThreads, 0, v->mf->ctx->stream>>>((double *)v->d, n, [a] __device__(double, long long) { return a; }); }
  else { float a = (float)value; map1_kernel<float><<<grid_for(n), kThreads, 0, v->mf->ctx->stream>>>((float *)v->d, n, [a] __device__(float, long long) { return a; }); }
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

int dgx_vec_scale(dgx_vec *v, double s) {
  DGX_REQUIRE(v, "null argument");
  const long long n = v->n_owned;
  if (v->dtype == DGX_F64) { double a = s; map1_kernel<double><<<grid_for(n), kThreads, 0, v->mf->ctx->stream>>>((double *)v->d, n, [a] __device__(double x, long long) { return a * x; }); }
  else { float a = (float)s; map1_kernel<float><<<grid_for(n), kThreads, 0, v->mf->ctx->stream>>>((float *)v->d, n, [a] __device__(float x, long long) { return a * x; }); }
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

}  // extern "C"

// dst = f(dst, src) with both same dtype
template <typename F64, typename F32>
static int map2_same(dgx_vec *dst, const dgx_vec *src, F64 f64, F32 f32) {
  DGX_REQUIRE(same_shape(dst, src) && dst->dtype == src->dtype, "vector shape / dtype mismatch");
  const long long n = dst->n_owned;
  cudaStream_t st = dst->mf->ctx->stream;
  if (dst->dtype == DGX_F64) map2_kernel<double, double><<<grid_for(n), kThreads, 0, st>>>((double *)dst->d, (const double *)src->d, n, f64);
  else map2_kernel<float, float><<<grid_for(n), kThreads, 0, st>>>((float *)dst->d, (const float *)src->d, n, f32);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

extern "C" {

int dgx_vec_copy(dgx_vec *dst, const dgx_vec *src) {
  DGX_REQUIRE(same_shape(dst, src), "vector shape mismatch");
  const long long n = dst->n_owned;
  cudaStream_t st = dst->mf->ctx->stream;
  if (dst->dtype == src->dtype) {
    DGX_CUDA_CHECK(cudaMemcpyAsync(dst->d, src->d, (size_t)n * dst->elem(), cudaMemcpyDeviceToDevice, st));
    return DGX_OK;
  }
  if (dst->dtype == DGX_F64) map2_kernel<double, float><<<grid_for(n), kThreads, 0, st>>>((double *)dst->d, (const float *)src->d, n, [] __device__(double, float s) { return (double)s; });
  else map2_kernel<float, double><<<grid_for(n), kThreads, 0, st>>>((float *)dst->d, (const double *)src->d, n, [] __device__(float, double s) { return (float)s; });
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

int dgx_vec_equ(dgx_vec *dst, double a, const dgx_vec *src) {
  const float af = (float)a;
  return map2_same(dst, src, [a] __device__(double, double s) { return a * s; }, [af] __device__(float, float s) { return af * s; });
}

int dgx_vec_add(dgx_vec *dst, double a, const dgx_vec *src) {
  const float af = (float)a;
  return map2_same(dst, src, [a] __device__(double d, double s) { return d + a * s; }, [af] __device__(float d, float s) { return d + af * s; });
}

int dgx_vec_sadd(dgx_vec *dst, double s_, double a, const dgx_vec *src) {
  const float af = (float)a, sf = (float)s_;
  return map2_same(dst, src, [a, s_] __device__(double d, double s) { return s_ * d + a * s; },
                   [af, sf] __device__(float d, float s) { return sf * d + af * s; });
}

int dgx_vec_scale_by(dgx_vec *dst, const dgx_vec *factors) {
  return map2_same(dst, factors, [] __device__(double d, double s) { return d * s; }, [] __device__(float d, float s) { return d * s; });
}

int dgx_vec_dot(const dgx_vec *a, const dgx_vec *b, double *out) {
  DGX_REQUIRE(same_shape(a, b) && a->dtype == b->dtype && out, "vector shape / dtype mismatch");
  return reduce_dispatch<0, false>(a->mf, a->dtype, a->d, b->d, nullptr, nullptr, 0.0, a->n_owned, out, false);
}

int dgx_vec_l2_norm(const dgx_vec *a, double *out) {
  double s = 0;
  int rc = dgx_vec_dot(a, a, &s);
  if (rc) return rc;
  *out = sqrt(s);
  return DGX_OK;
}

int dgx_vec_linfty_norm(const dgx_vec *a, double *out) {
  DGX_REQUIRE(a && out, "null argument");
  return reduce_dispatch<1, false>(a->mf, a->dtype, a->d, nullptr, nullptr, nullptr, 0.0, a->n_owned, out, true);
}

int dgx_vec_mean_value(const dgx_vec *a, double *out) {
  DGX_REQUIRE(a && out, "null argument");
  double s = 0;
  int rc = reduce_dispatch<2, false>(a->mf, a->dtype, a->d, nullptr, nullptr, nullptr, 0.0, a->n_owned, &s, false);
  if (rc) return rc;
  *out = s / (double)a->n_global;
  return DGX_OK;
}

int dgx_vec_add_and_dot(dgx_vec *dst, double a, const dgx_vec *v, const dgx_vec *w, double *out) {
  DGX_REQUIRE(same_shape(dst, v) && same_shape(dst, w) && dst->dtype == v->dtype && dst->dtype == w->dtype && out, "vector mismatch");
  return reduce_dispatch<0, true>(dst->mf, dst->dtype, nullptr, w->d, dst->d, v->d, a, dst->n_owned, out, false);
}

// x += a p ; r += b v ; *out = r . r  -- the two vector updates and the residual norm of one SolverCG iteration, one pass
int dgx_vec_cg_update(dgx_vec *x, double a, const dgx_vec *p, dgx_vec *r, double b, const dgx_vec *v, double *out) {
  DGX_REQUIRE(same_shape(x, p) && same_shape(x, r) && same_shape(x, v) && x->dtype == p->dtype && x->dtype == r->dtype && x->dtype == v->dtype && out,
              "vector mismatch");
  return reduce_dispatch<0, true>(x->mf, x->dtype, nullptr, r->d, r->d, v->d, b, x->n_owned, out, false, x->d, p->d, a);
}

// interpolate_velocity (NS.cc:2403-2418) fused into one pass: dst = a * x - (b * y), with the reference's rounding
// order (equ, equ, -=): a * x and b * y are rounded separately before the subtraction.
int dgx_vec_interpolate(dgx_vec *dst, double a, const dgx_vec *x, double b, const dgx_vec *y) {
  DGX_REQUIRE(same_shape(dst, x) && same_shape(dst, y) && dst->dtype == x->dtype && dst->dtype == y->dtype, "vector shape / dtype mismatch");
  const long long n = dst->n_owned;
  cudaStream_t st = dst->mf->ctx->stream;
  if (n == 0) return DGX_OK;
  if (dst->dtype == DGX_F64) {
    const double *xp = (const double *)x->d, *yp = (const double *)y->d;
    map1_kernel<double><<<grid_for(n), kThreads, 0, st>>>((double *)dst->d, n, [=] __device__(double, long long i) { return __dsub_rn(__dmul_rn(a, xp[i]), __dmul_rn(b, yp[i])); });
  } else {
    const float *xp = (const float *)x->d, *yp = (const float *)y->d;
    const float af = (float)a, bf = (float)b;
    map1_kernel<float><<<grid_for(n), kThreads, 0, st>>>((float *)dst->d, n, [=] __device__(float, long long i) { return __fsub_rn(__fmul_rn(af, xp[i]), __fmul_rn(bf, yp[i])); });
  }
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

int dgx_vec_update_ghost_values(dgx_vec *v) {
  DGX_REQUIRE(v, "null argument");
  return ghost_exchange(v);
}

int dgx_vec_zero_out_ghost_values(dgx_vec *v) {
  DGX_REQUIRE(v, "null argument");
  if (v->n_ghost)
    DGX_CUDA_CHECK(cudaMemsetAsync((char *)v->d + (size_t)v->n_owned * v->elem(), 0, (size_t)v->n_ghost * v->elem(), v->mf->ctx->stream));
  return DGX_OK;
}

int dgx_vec_compress_add(dgx_vec *v) {
  DGX_REQUIRE(v, "null argument");
  return DGX_OK;
}


// One Gram-Schmidt sweep of SolverGMRES: h_j = w . V_j for j < k (one pass over w, one all-reduce, one host read-back),
// then w -= sum_j h_j V_j (one pass).  k <= 32.
int dgx_vec_orthogonalize(dgx_vec *w, dgx_vec *const *V, int k, double *h) { return dgx_vec_orthogonalize_checked(w, V, k, h, nullptr); }

// same sweep; *w_dot_w (if not NULL) receives w . w BEFORE the sweep, reduced in the same pass (the reference's loss-of-
// orthogonality test compares the norms before and after, Kelley 1995 as used by deal.II's SolverGMRES)
int dgx_vec_orthogonalize_checked(dgx_vec *w, dgx_vec *const *V, int k, double *h, double *w_dot_w) {
  DGX_REQUIRE(w && V && h && k >= 1 && k + (w_dot_w ? 1 : 0) <= kMaxMulti, "orthogonalize: 1 <= k <= 32 (31 with the norm)");
  dgx_ctx *ctx = w->mf->ctx;
  const long long n = w->n_owned;
  const int grid = grid_for(n, 8);
  constexpr int KB = 8;
  DGX_REQUIRE((size_t)grid * KB + kScratchPartial <= ctx->scratch_doubles, "reduction scratch too small");
  double *result = ctx->d_scratch, *partial = ctx->d_scratch + kScratchPartial;
  static_assert(kMaxMulti <= 32, "result slots of the context scratch");
  auto run = [&](auto tag) -> int {
    using T = decltype(tag);
    PtrTable<T> tab;
    for (int j = 0; j < k; ++j) {
      DGX_REQUIRE(V[j] && same_shape(w, V[j]) && V[j]->dtype == w->dtype, "basis vector shape / dtype mismatch");
      tab.v[j] = (const T *)V[j]->d;
      tab.h[j] = 0;
    }
    const int kd = k + (w_dot_w ? 1 : 0);   // dot products of this pass: the basis, then w itself
    if (w_dot_w) { tab.v[k] = (const T *)w->d; tab.h[k] = 0; }
    for (int k0 = 0; k0 < kd; k0 += KB) {
      multi_dot_kernel<T, KB><<<grid, kThreads, 0, ctx->stream>>>((const T *)w->d, tab, k0, kd, n, partial, result);
      count_launch();
    }
    DGX_CUDA_CHECK(cudaGetLastError());
    if (!w->mf->replicated) {
      int rc = allreduce_sum(ctx, result, kd);
      if (rc) return rc;
    }
    DGX_CUDA_CHECK(cudaMemcpyAsync(ctx->h_scratch, result, kd * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < k; ++j) { h[j] = ctx->h_scratch[j]; tab.h[j] = h[j]; }
    if (w_dot_w) *w_dot_w = ctx->h_scratch[k];
    if (n) { multi_axpy_kernel<T><<<grid_for(n), kThreads, 0, ctx->stream>>>((T *)w->d, tab, k, n); count_launch(); }
    DGX_CUDA_CHECK(cudaGetLastError());
    return DGX_OK;
  };
  return w->dtype == DGX_F64 ? run(double()) : run(float());
}

}  // extern "C"

