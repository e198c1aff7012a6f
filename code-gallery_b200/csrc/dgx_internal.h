// Internal declarations of libdgx (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dgx.h"

namespace dgx {

void set_error(const std::string &msg);
extern long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

#define DGX_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      ::dgx::set_error(std::string(#expr) + ": " + cudaGetErrorString(err__));                \
      return DGX_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define DGX_REQUIRE(cond, msg)                                                                \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::dgx::set_error(std::string(msg) + " (" #cond ")");                                    \
      return DGX_ERR_INVALID;                                                                 \
    }                                                                                         \
  } while (0)

// Reduction scratch of a context (dgx_ctx::d_scratch, doubles): [0, 32) results (up to 32 dot products of one Gram-Schmidt
// sweep), [32] last-block ticket of the two-level reductions (per context: two contexts on one device do not share it),
// [kScratchPartial, ...) per-block partial sums.
constexpr int kScratchPartial = 34;
#ifdef __CUDACC__
__device__ __forceinline__ unsigned int *scratch_ticket(double *partial) { return reinterpret_cast<unsigned int *>(partial - 2); }
#endif

constexpr int MAX_DIM = 3;
constexpr int MAX_N = 9;  // degree <= 8

// ---- 1-D tables (host, long double -> double) -------------------------------------------------
struct Shape1D {
  int n = 0, nq = 0;
  std::vector<double> nodes, xq, wq;
  std::vector<double> S, D;         // [nq][n]
  std::vector<double> f[2], g[2];   // l_i(s), l_i'(s)
  std::vector<double> M, K;         // [n][n] sum_q w S S, sum_q w D D
  std::vector<double> Minv;         // [n][n]
  std::vector<double> Kt;           // Minv K
  std::vector<double> a[2], b[2];   // Minv f_s, Minv g_s
  std::vector<double> P[2];         // [n][n] P_c[j][i] = l_i((x_j + c)/2)
  // collocation (Lagrange basis on the Gauss points) tables, only when nq == n:
  std::vector<double> Dhat;         // [nq][nq] derivative of the Gauss-point Lagrange basis at the Gauss points (= D S^-1)
  std::vector<double> fhat[2], ghat[2];   // value / derivative of that basis at the end points (= f S^-1, g S^-1)
};
const Shape1D &get_shape(int degree, int nq);

// ---- mesh -------------------------------------------------------------------------------------
struct Mesh {
  int dim = 0;
  int cells0[MAX_DIM] = {1, 1, 1};
  int n_refine = 0;
  double lo[MAX_DIM] = {0, 0, 0}, hi[MAX_DIM] = {1, 1, 1};
  bool periodic[MAX_DIM] = {false, false, false};
  int bids[2 * MAX_DIM] = {0, 1, 2, 3, 4, 5};
  // vertex positions per level (dgx_mesh_set_vertices): lattice of (cells_dir + 1)^dim points, x fastest, dim doubles each; empty =
  // the Cartesian box.  A level with vertices is a MappingQ1 mesh: operators use the per-quadrature-point metric path.
  std::vector<std::vector<double>> vertices;
  bool general(int level) const { return level >= 0 && level < (int)vertices.size() && !vertices[level].empty(); }

  long long n_cells(int level) const;
  int cells_dir(int level, int d) const { return cells0[d] << level; }
  double h(int level, int d) const { return (hi[d] - lo[d]) / double(cells_dir(level, d)); }
  void coords_of(int level, long long idx, int *c) const;
  long long index_of(int level, const int *c) const;
  // neighbour across face f (2d+s); returns cell index or -1-boundary_id
  long long neighbor(int level, long long idx, int f) const;
};

}  // namespace dgx

struct dgx_mesh {
  dgx::Mesh m;
};

namespace dgx { struct P2PScalar; }

struct dgx_ctx {
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // ghost exchange overlapped with interior work: pack + NCCL run on comm_stream between these two events
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_src_ready = nullptr, ev_ghosts_done = nullptr;
  void *nccl_comm = nullptr;
  double *d_scratch = nullptr;   // reduction scratch (device)
  double *h_scratch = nullptr;   // pinned
  size_t scratch_doubles = 0;
  bool p2p_off = false;          // CUDA IPC / peer access not available: ghost imports go through NCCL
  dgx::P2PScalar *p2ps = nullptr;   // all-reduce of a few doubles over peer memory (p2p.cu)
  bool p2ps_tried = false;
};

namespace dgx { struct P2PBox; }

struct dgx_mf {
  dgx_ctx *ctx = nullptr;
  dgx_mesh *mesh = nullptr;
  int level = 0;
  int dtype = DGX_F64;
  int dim = 0;
  long long first_cell = 0, n_owned = 0, n_ghost = 0, n_global = 0;
  double h[dgx::MAX_DIM] = {1, 1, 1};
  std::vector<int> nbr_host;            // [2*dim][n_owned]
  std::vector<long long> ghost_global;  // [n_ghost]
  int *d_nbr = nullptr;
  bool all_interior = false;            // no boundary faces at all (periodic everywhere)
  // Replicated level (multigrid levels too small to be worth distributing): every rank holds ALL cells, computes the same
  // numbers redundantly and never communicates on this level (no ghosts, reductions stay local).
  bool replicated = false;
  // owned cells == the box [box_lo, box_lo + box_ext) of cell coordinates (true for Morton-aligned partitions)
  bool box_ok = false;
  int box_lo[dgx::MAX_DIM] = {0, 0, 0}, box_ext[dgx::MAX_DIM] = {1, 1, 1};
  // ghost exchange plan: per peer rank, the local owned cells to send and the ghost slot range to receive
  struct Peer {
    int rank;
    std::vector<int> send_cells;        // local indices
    long long recv_first = 0, recv_count = 0;  // ghost slots [recv_first, recv_first+recv_count)
    int *d_send_cells = nullptr;
  };
  std::vector<Peer> peers;
  // marching kernel job lists (tile columns that need no ghost data / that do), built on first use for jobs_NZ
  int *d_jobs = nullptr;
  int n_jobs_interior = 0, n_jobs_boundary = 0, jobs_NZ = 0;
  int *d_jobs_async = nullptr;          // all tile columns, those that read ghost cells from their first plane on last
  int jobs_async_NZ = 0;
  // ghost import over NVLink peer memory (p2p.cu): one mailbox per cell size in bytes, created at the first exchange
  std::map<size_t, dgx::P2PBox *> p2p;
  // boundary faces (local cell, face number) per boundary id, built on first use (diagnostics.cu: lift and drag)
  struct BFaces { int *d = nullptr; int n = 0; };
  std::map<int, BFaces> bfaces;
  // non-Cartesian level (Mesh::general): metric tables of the owned cells per number of 1-D quadrature points (pressure_general.cu)
  bool general = false;
  struct GeneralMetric { void *cellG = nullptr, *faceA = nullptr; size_t bytes = 0; };
  std::map<int, GeneralMetric> gmetric;
  double *d_cellX = nullptr;            // vertices of the owned cells [cell][2^dim][dim] (diagnostics.cu), built on first use
  std::vector<double> h_cellX;
};

struct dgx_vec {
  dgx_mf *mf = nullptr;
  int degree = 0, ncomp = 1, dtype = DGX_F64;
  long long dofs_per_cell = 0, n_owned = 0, n_ghost = 0, n_global = 0;
  void *d = nullptr;        // [n_owned + n_ghost]
  void *d_sendbuf = nullptr;
  size_t sendbuf_bytes = 0;
  bool borrowed = false;    // owned by an operator (inverse diagonal)
  size_t elem() const { return dtype == DGX_F64 ? 8 : 4; }
};

struct dgx_op {
  dgx_mf *mf = nullptr;
  int kind = DGX_OP_PRESSURE;
  int degree = 1, ncomp = 1;
  double Re = 1.0, dt = 1.0;
  int stage = 1;
  int variant = 0;
  dgx_vec *u_extr = nullptr;
  dgx_vec *inv_diag = nullptr;
  dgx_vec *h_src = nullptr, *h_dst = nullptr;   // staging vectors of the host entry point
  const dgx_vec *post_scale = nullptr;          // velocity operator only, set around one vmult: dst = post_scale .* (A src)
  // pipelined host entry point: upload segments (cells [b, e), in upload order) and compute groups (cells [b, e), computable
  // once upload segment `need` is in; in the order they become computable)
  struct HostSeg { long long b, e; int need; };
  std::vector<HostSeg> hp_up, hp_grp;
  int hp_lg = 0;                                // log2(cells per chunk); 0 = not decided yet, -1 = pipeline not used
  std::vector<cudaEvent_t> hp_in, hp_done;
  cudaStream_t hp_s_in = nullptr, hp_s_out = nullptr;
  // several ranks: the cells other ranks need (union of the send lists) go first -- gathered on the host into a pinned staging buffer,
  // uploaded in one copy, scattered into the vector and exchanged -- so that the pipeline below never waits for a neighbour rank
  std::vector<int> hp_bcells;
  int *hp_d_bcells = nullptr;
  void *hp_stage_h = nullptr, *hp_stage_d = nullptr;
  cudaEvent_t hp_ev_ghosts = nullptr;
  // probed inverse diagonals of a non-Cartesian level per (TR_BDF2 stage, dt): the geometry does not change between time steps, so
  // the n^dim probing applications of compute_diagonal run once per pair (pressure_general.cu)
  struct DiagCache { int stage; double dt; dgx_vec *v; };
  std::vector<DiagCache> gen_diag_cache;
  dgx_vec *work[4] = {nullptr, nullptr, nullptr, nullptr};   // Krylov work vectors, allocated on first use and kept
  std::vector<dgx_vec *> pool;                                // further work vectors (GMRES basis, diagonal probes), kept
};

namespace dgx {
// time-stepping constants, bit-identical to NS.cc:396-397, 286-287
double gamma_trbdf2();
// kernel launchers (defined in .cu files); all enqueue on mf->ctx->stream
int pressure_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add);
int pressure_apply_fast_overlapped(dgx_op *op, dgx_vec *dst, dgx_vec *src, bool add);
int pressure_diagonal(dgx_op *op, dgx_vec *diag_inv);
// fused operator + vector update: dst = a x + b xold + c d .* (rhs - A x)   (xold, d may be null)
struct FusedUpdate {
  const dgx_vec *xold = nullptr, *d = nullptr, *rhs = nullptr;
  double a = 0, b = 0, c = 0;
};
int pressure_apply_fused(dgx_op *op, dgx_vec *dst, dgx_vec *x, const FusedUpdate &fu);
// pressure operator on a level with vertex positions (MappingQ1 metric at every quadrature point); bfac scales the boundary penalty
int pressure_general_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, const FusedUpdate *fu, double bfac);
int pressure_general_diagonal(dgx_op *op, dgx_vec *diag_inv);
int pressure_general_coarse_cg(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info,
                               double *d_res, int *d_count, int log_cap);
int mass_general_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse);
void general_metric_destroy(dgx_mf *mf);
int general_metric_host_tables(const dgx_mf *mf, int n, std::vector<double> &cellG, std::vector<double> &faceA);
long long general_metric_bytes(dgx_mf *mf, int n);
// i-th pooled work vector of the operator's space (created on first use, lives as long as the operator)
int op_pool_vec(dgx_op *op, int i, dgx_vec **out);
int pressure_coarse_cg(dgx_op *op, dgx_vec *x, const dgx_vec *b, dgx_vec *r, dgx_vec *p, dgx_vec *v, const double *d_tol, int max_it, int *d_info, double *d_res,
                       int *d_count, int log_cap);
int mass_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add);
int mass_apply_any(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add, bool inverse);
int mass_inverse_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src);
int mass_cg_front(dgx_op *op, dgx_vec *p, const dgx_vec *r, double beta, bool first, dgx_vec *v, double *pv);
int velocity_apply(dgx_op *op, dgx_vec *dst, const dgx_vec *src, bool add);
int velocity_diagonal(dgx_op *op, dgx_vec *diag_inv);
int ghost_exchange(dgx_vec *v);
int scatter_cells(dgx_vec *v, const void *d_staged, const int *d_cells, long long ncells, cudaStream_t stream);
// Ghost cells of an import that is still in flight (p2p.cu, asynchronous mode): the consumer kernel waits for the neighbours'
// sequence numbers itself and reads the ghost cells straight from the mailbox slot.  base == nullptr: ghosts are in the vector tail.
struct GhostSrc {
  const char *base = nullptr;                  // ghost cells of this exchange in ghost-slot order
  const unsigned long long *flags = nullptr;   // [nranks] arrival sequence numbers, written by the neighbours' copy engines
  const int *peer_ranks = nullptr;             // [npeers]
  int npeers = 0;
  unsigned long long seq = 0;
};
// pack on the context's stream, then copy-engine transfers into the neighbours' mailboxes + sequence numbers on the communication
// stream; *done = 0: not available (caller imports synchronously)
int p2p_ghost_import_async(dgx_vec *v, GhostSrc *gs, int *done);
int pressure_apply_fast_async(dgx_op *op, dgx_vec *dst, dgx_vec *src, bool add);
// ghost import over peer memory: *done = 1 when it was enqueued on `stream`, 0 when the caller must use NCCL
int p2p_ghost_exchange(dgx_vec *v, cudaStream_t stream, int *done);
void p2p_destroy(dgx_mf *mf);
int p2p_allreduce(dgx_ctx *ctx, double *dvals, int n, bool is_max, int *done);
void p2p_scalar_destroy(dgx_ctx *ctx);
int nccl_allreduce(dgx_ctx *ctx, double *dvals, int n, bool is_max);
// same exchange on the context's communication stream: starts when the work already queued on the compute stream is
// done; ghost_exchange_wait makes the compute stream wait for the ghosts
int ghost_exchange_start(dgx_vec *v);
// pressure operator on the owned cells [cell_begin, cell_end) only, on `stream` (single rank)
int pressure_apply_range(dgx_op *op, dgx_vec *dst, const dgx_vec *src, long long cell_begin, long long cell_end, cudaStream_t stream);
int ghost_exchange_wait(dgx_ctx *ctx);
// internal form of dgx_mf_create: replicated = every rank owns the whole level
int mf_create(dgx_ctx *ctx, dgx_mesh *mesh, int level, int dtype, bool replicated, dgx_mf **out);
// in-place all-gather of equal per-rank pieces: rank r's piece sits at buf + r * bytes_per_rank already
int allgather_inplace(dgx_ctx *ctx, void *buf, size_t bytes_per_rank);
int allreduce_sum(dgx_ctx *ctx, double *vals, int n);
int allreduce_max(dgx_ctx *ctx, double *vals, int n);
int comm_init(dgx_ctx *ctx, const void *id);
int comm_destroy(dgx_ctx *ctx);
int comm_unique_id(void *out);
}  // namespace dgx
