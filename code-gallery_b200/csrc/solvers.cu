// Device-resident Krylov / Chebyshev / multigrid layer: the deal.II solver objects the reference wires around
// the operator in diffusion_step / projection_step / project_grad (NS.cc:2424-2577), restated for the device
// vector.  Algebra follows deal.II exactly so that iteration counts can be compared:
//   SolverCG                 -> dgx_solve_cg            (NS.cc:2487, 2542/2546; mass solve NS.cc:2575-2576)
//   SolverGMRES              -> dgx_solve_gmres         (NS.cc:2451, 2463)
//   PreconditionChebyshev    -> dgx_cheb_*              (NS.cc:2493-2516)
//   MGTransferMatrixFree     -> dgx_transfer_*          (NS.cc:2490-2491)
//   Multigrid + PreconditionMG + MGCoarseGridIterativeSolver -> dgx_mg_* (NS.cc:2519-2537)
// The host only sequences kernel launches and reads back the few scalars (dot products) the algorithms
// branch on; all vectors and all arithmetic on them stay on the device.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "pressure_op.cuh"

using namespace dgx;

struct dgx_cheb {
  dgx_op *op = nullptr;
  int degree = 3;
  double smoothing_range = 15.0;
  int eig_cg_n_iterations = 10;
  bool initialized = false;
  double theta = 1.0, delta = 0.0, max_eig_est = 0.0, min_eig_est = 0.0;
  dgx_vec *t = nullptr, *x_old = nullptr, *x_tmp = nullptr;   // work vectors
};

struct dgx_mg {
  dgx_ctx *ctx = nullptr;
  dgx_mesh *mesh = nullptr;
  int degree = 1, n_levels = 0;
  double Re = 1.0, dt = 1.0;
  int stage = 1;
  std::vector<dgx_mf *> mf;
  std::vector<dgx_op *> op;
  std::vector<dgx_cheb *> smoother;          // [level], level 0 unused (coarse CG)
  std::vector<dgx_vec *> defect, sol, t;     // [level]
  dgx_mf *mf_fine64 = nullptr;               // not owned: the caller's fp64 MatrixFree (for shape checks)
  double coarse_tol = 0.0;
  int coarse_max_it = 10000;
  std::vector<int> coarse_its;               // iterations of every coarse solve of the last outer solve
  // coarse CG work vectors
  dgx_vec *cg_r = nullptr, *cg_p = nullptr, *cg_v = nullptr;
  // device-side log of the single-CTA coarse solves: ring of (iterations, converged) pairs and residuals, appended by the
  // kernel itself through d_count (so that a replayed CUDA graph logs too); the tolerance is read from d_tol
  static constexpr int kLog = 8192;
  int *d_info = nullptr, *d_count = nullptr;
  double *d_res = nullptr, *d_tol = nullptr;
  long long log_seen = 0;                    // entries already checked for convergence
  long long log_setup = 0;                   // log position at the last setup (dgx_mg_coarse_iterations reports from here)
  // the V-cycle as a CUDA graph per (stage, dt, estimate generation): launch sequence and kernel arguments are identical from
  // one application to the next (defect / solution / work vectors are owned by this object)
  // A combination is captured only when it comes back in a LATER solve (capturing costs about one eager cycle on one GPU and
  // several with NCCL nodes): with recomputed estimates every solve has its own generation and stays eager.
  struct CycleGraph { int stage; double dt; unsigned long long gen; long long first_setup = 0; cudaGraphExec_t exec = nullptr; long long launches = 0; };
  long long n_setups = 0;
  std::vector<CycleGraph> graphs;
  unsigned long long est_gen = 0;            // bumped whenever smoother estimates are recomputed (they are kernel arguments)
  bool graphs_ok = true, warmed = false;
  // Chebyshev eigenvalue estimates per (TR_BDF2 stage, dt): they depend on nothing else (the level operators are fixed by the
  // mesh, kappa(stage, dt) and the degree), so a time loop pays the 10 CG iterations per level once per stage, not per solve
  struct EigKey { int stage; double dt; };
  std::vector<std::pair<EigKey, std::vector<std::array<double, 4>>>> eig_cache;
  bool cache_estimates = true;
  int last_setup_cached = 0;
};

extern "C" int mg_check_coarse_log(dgx_mg *mg);
extern "C" int mg_cycle(dgx_mg *mg);

namespace {

__global__ void set_scalar_kernel(double *p, double v) { *p = v; }

constexpr int kThreads = 256;
inline int grid_for(long long n, int per_thread = 4) {
  long long b = (n + (long long)kThreads * per_thread - 1) / ((long long)kThreads * per_thread);
  return (int)std::max<long long>(1, std::min<long long>(b, 148 * 8));
}

// out = a * x + b * xold + c * d .* (rhs - t)   (null pointers drop their term)
template <typename T>
__global__ void __launch_bounds__(kThreads) cheb_update_kernel(T *__restrict__ out, const T *x, const T *xold, const T *__restrict__ d,
                                                               const T *__restrict__ rhs, const T *t, T a, T b, T c, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    T r = rhs[i];
    if (t) r -= t[i];
    T v = c * d[i] * r;
    if (x) v += a * x[i];
    if (xold) v += b * xold[i];
    out[i] = v;
  }
}

// v[i] = (global index) % 11   (PreconditionChebyshev::set_initial_guess)
template <typename T>
__global__ void iota_mod11_kernel(T *v, long long first, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[i] = (T)((first + i) % 11);
}

template <typename T>
__global__ void add_scalar_kernel(T *v, T a, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] += a;
}

// ---- MGTransferMatrixFree for FE_DGQ: child = (P_cx (x) P_cy (x) P_cz) parent ---------------------------
template <typename T, int n>
struct TransferParams {
  T P[2][n * n];   // P[c][j*n+i] = l_i((x_j + c)/2)
};

// one block per parent cell; thread = (child, line).  PROLONG: fine[child] += P parent;  else coarse += P^T fine[child]
template <int dim, int n, typename T, bool PROLONG>
__global__ void __launch_bounds__((1 << dim) * ipow(n, dim - 1))
transfer_kernel(const __grid_constant__ TransferParams<T, n> P, long long n_parents, int ncomp, T *__restrict__ dst, const T *__restrict__ src) {
  constexpr int NC = 1 << dim, NL = ipow(n, dim - 1), ND = NL * n, NT = NC * NL;
  __shared__ T buf[NC][ND];
  __shared__ T par[ND];
  // one block per (parent cell, component): cells are component-blocked, [cell][comp][n^dim]
  const long long parent = blockIdx.x / ncomp;
  const int comp = (int)(blockIdx.x - parent * ncomp);
  const size_t coarse_off = ((size_t)parent * ncomp + comp) * ND;
  auto fine_off = [&](int child) { return (((size_t)parent * NC + child) * ncomp + comp) * ND; };
  const int tid = threadIdx.x, c = tid / NL, l = tid % NL;
  if (PROLONG) {
    for (int i = tid; i < ND; i += NT) par[i] = src[coarse_off + i];
  } else {
    for (int i = tid; i < NC * ND; i += NT) buf[i / ND][i % ND] = src[fine_off(i / ND) + i % ND];
  }
  __syncthreads();
  if (PROLONG)
    for (int i = l; i < ND; i += NL) buf[c][i] = par[i];
  __syncthreads();
#pragma unroll
  for (int d = 0; d < dim; ++d) {
    int base, stride;
    if (d == 0) { base = l * n; stride = 1; }
    else if (d == 1 && dim == 3) { base = (l % n) + n * n * (l / n); stride = n; }
    else if (d == 1) { base = l; stride = n; }
    else { base = l; stride = n * n; }
    const int cd = (c >> d) & 1;
    T u[n];
#pragma unroll
    for (int i = 0; i < n; ++i) u[i] = buf[c][base + stride * i];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      T acc = 0;
#pragma unroll
      for (int i = 0; i < n; ++i) acc += (PROLONG ? P.P[cd][j * n + i] : P.P[cd][i * n + j]) * u[i];
      buf[c][base + stride * j] = acc;
    }
    __syncthreads();
  }
  if (PROLONG) {
    for (int i = tid; i < NC * ND; i += NT) dst[fine_off(i / ND) + i % ND] += buf[i / ND][i % ND];
  } else {
    for (int i = tid; i < ND; i += NT) {
      T s = 0;
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) s += buf[cc][i];
      dst[coarse_off + i] += s;
    }
  }
}

template <int dim, int n, typename T>
int launch_transfer(dgx_ctx *ctx, bool prolong, long long n_parents, int ncomp, void *dst, const void *src) {
  const Shape1D &s = get_shape(n - 1, n);
  TransferParams<T, n> P;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < n * n; ++i) P.P[c][i] = (T)s.P[c][i];
  constexpr int NT = (1 << dim) * ipow(n, dim - 1);
  if (n_parents == 0) return DGX_OK;
  if (prolong) transfer_kernel<dim, n, T, true><<<(unsigned)(n_parents * ncomp), NT, 0, ctx->stream>>>(P, n_parents, ncomp, (T *)dst, (const T *)src);
  else transfer_kernel<dim, n, T, false><<<(unsigned)(n_parents * ncomp), NT, 0, ctx->stream>>>(P, n_parents, ncomp, (T *)dst, (const T *)src);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

template <int dim, typename T>
int transfer_by_degree(int degree, dgx_ctx *ctx, bool prolong, long long np, int ncomp, void *dst, const void *src) {
  switch (degree) {
    case 1: return launch_transfer<dim, 2, T>(ctx, prolong, np, ncomp, dst, src);
    case 2: return launch_transfer<dim, 3, T>(ctx, prolong, np, ncomp, dst, src);
    case 3: return launch_transfer<dim, 4, T>(ctx, prolong, np, ncomp, dst, src);
    case 4: return launch_transfer<dim, 5, T>(ctx, prolong, np, ncomp, dst, src);
    case 5: return launch_transfer<dim, 6, T>(ctx, prolong, np, ncomp, dst, src);
    case 6: return launch_transfer<dim, 7, T>(ctx, prolong, np, ncomp, dst, src);
    default: set_error("transfer: degree must be in 1..6"); return DGX_ERR_UNSUPPORTED;
  }
}

int transfer(bool prolong, dgx_vec *dst, const dgx_vec *src) {
  DGX_REQUIRE(dst && src, "null argument");
  const dgx_vec *fine = prolong ? dst : src, *coarse = prolong ? src : dst;
  DGX_REQUIRE(fine->mf->level == coarse->mf->level + 1 && fine->mf->mesh == coarse->mf->mesh, "vectors must live on consecutive levels of one mesh");
  DGX_REQUIRE(fine->degree == coarse->degree && fine->ncomp == coarse->ncomp && fine->dtype == coarse->dtype, "transfer: spaces of equal degree, components and number type");
  const int ncomp = fine->ncomp;
  const int dim = fine->mf->dim, nc = 1 << dim;
  // the parents of this rank's fine cells: the coarse cells [par_first, par_first + np) in global numbering
  if (fine->mf->first_cell % nc != 0 || fine->mf->n_owned % nc != 0) {
    set_error("transfer: the owned fine cells must be whole families of children");
    return DGX_ERR_UNSUPPORTED;
  }
  const long long par_first = fine->mf->first_cell / nc, np = fine->mf->n_owned / nc;
  // coarse side: distributed with the same cut (parents owned here), or replicated (every rank holds all coarse cells)
  const bool gather = coarse->mf->replicated && !fine->mf->replicated;
  if (!gather && (coarse->mf->first_cell != par_first || coarse->mf->n_owned != np)) {
    set_error("transfer: children of an owned parent cell must be owned by the same rank");
    return DGX_ERR_UNSUPPORTED;
  }
  dgx_ctx *ctx = fine->mf->ctx;
  const long long coarse_off = (par_first - coarse->mf->first_cell) * coarse->dofs_per_cell;   // 0 unless the coarse level is replicated
  const size_t es = coarse->elem();
  void *dptr = prolong ? dst->d : (char *)dst->d + (size_t)coarse_off * es;
  const void *sptr = prolong ? (const char *)src->d + (size_t)coarse_off * es : src->d;
  int rc;
  if (dim == 2) rc = fine->dtype == DGX_F64 ? transfer_by_degree<2, double>(fine->degree, ctx, prolong, np, ncomp, dptr, sptr)
                                            : transfer_by_degree<2, float>(fine->degree, ctx, prolong, np, ncomp, dptr, sptr);
  else rc = fine->dtype == DGX_F64 ? transfer_by_degree<3, double>(fine->degree, ctx, prolong, np, ncomp, dptr, sptr)
                                   : transfer_by_degree<3, float>(fine->degree, ctx, prolong, np, ncomp, dptr, sptr);
  if (rc || prolong || !gather) return rc;
  // restriction into a replicated level: every rank has filled the piece of its own parents; one all-gather completes the
  // vector on all ranks (equal pieces: the fine level is cut into equal chunks).  Nothing is sent on the way up.
  DGX_REQUIRE(np * ctx->nranks == coarse->mf->n_owned && par_first == np * ctx->rank, "transfer: unequal pieces in the gather to a replicated level");
  return allgather_inplace(ctx, dst->d, (size_t)np * coarse->dofs_per_cell * es);
}

int cheb_update(dgx_vec *out, const dgx_vec *x, const dgx_vec *xold, const dgx_vec *d, const dgx_vec *rhs, const dgx_vec *t, double a, double b, double c) {
  const long long n = out->n_owned;
  cudaStream_t st = out->mf->ctx->stream;
  if (n == 0) return DGX_OK;
  if (out->dtype == DGX_F64)
    cheb_update_kernel<double><<<grid_for(n), kThreads, 0, st>>>((double *)out->d, x ? (const double *)x->d : nullptr, xold ? (const double *)xold->d : nullptr,
                                                                  (const double *)d->d, (const double *)rhs->d, t ? (const double *)t->d : nullptr, a, b, c, n);
  else
    cheb_update_kernel<float><<<grid_for(n), kThreads, 0, st>>>((float *)out->d, x ? (const float *)x->d : nullptr, xold ? (const float *)xold->d : nullptr,
                                                                 (const float *)d->d, (const float *)rhs->d, t ? (const float *)t->d : nullptr, (float)a, (float)b, (float)c, n);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

// eigenvalues of a symmetric tridiagonal matrix (n <= ~30): cyclic Jacobi on the dense form
void tridiag_eigenvalues(const std::vector<double> &diag, const std::vector<double> &off, std::vector<double> &ev) {
  const int n = (int)diag.size();
  std::vector<double> A((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) A[i * n + i] = diag[i];
  for (int i = 0; i + 1 < n; ++i) A[i * n + i + 1] = A[(i + 1) * n + i] = off[i];
  for (int sweep = 0; sweep < 100; ++sweep) {
    double offn = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) offn += A[i * n + j] * A[i * n + j];
    if (offn < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double tau = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
        const double c = 1 / std::sqrt(1 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
      }
  }
  ev.resize(n);
  for (int i = 0; i < n; ++i) ev[i] = A[i * n + i];
  std::sort(ev.begin(), ev.end());
}

struct Precond {
  int kind = 0;            // 0 identity, 1 Jacobi (inverse diagonal of `op`), 2 multigrid, 3 inverse diagonal vector
  dgx_op *op = nullptr;
  dgx_mg *mg = nullptr;
  dgx_vec *diag = nullptr;
};

int apply_precond(const Precond &pc, dgx_vec *dst, dgx_vec *src) {
  switch (pc.kind) {
    case 0: return dgx_vec_copy(dst, src);
    case 1:
      if (!pc.op->inv_diag) { set_error("compute_diagonal() has not been called"); return DGX_ERR_INVALID; }
      return cheb_update(dst, nullptr, nullptr, pc.op->inv_diag, src, nullptr, 0.0, 0.0, 1.0);
    case 2: return dgx_mg_vmult(pc.mg, dst, src);
    case 3: return cheb_update(dst, nullptr, nullptr, pc.diag, src, nullptr, 0.0, 0.0, 1.0);   // dst = diag .* src in one pass
    default: set_error("unknown preconditioner kind"); return DGX_ERR_INVALID;
  }
}

#define RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

// Krylov work vectors live with the operator: cudaMalloc / cudaFree (which synchronises the device) stay out of the solves
int op_work(dgx_op *op, dgx_vec **w, int count) {
  for (int i = 0; i < count; ++i) {
    if (!op->work[i]) { RC(dgx_vec_create(op->mf, op->degree, op->ncomp, &op->work[i])); op->work[i]->borrowed = true; }
    w[i] = op->work[i];
  }
  return DGX_OK;
}

// deal.II SolverCG.  work: r, p, v, z.  Optional Lanczos coefficients for eigenvalue estimates.
int cg_core(dgx_op *A, dgx_vec *x, dgx_vec *b, const Precond &pc, int max_it, double tol, dgx_vec *r, dgx_vec *p, dgx_vec *v, dgx_vec *z,
            int *iters, double *final_res, std::vector<double> *lz_diag, std::vector<double> *lz_off, bool throw_on_fail) {
  double res = 0, r_dot_z = 0, prev_rz = 0, alpha = 0, beta = 0, prev_alpha = 0, eigen_beta_alpha = 0, pv = 0, rr = 0;
  RC(dgx_op_vmult(A, r, x));
  RC(dgx_vec_sadd(r, -1.0, 1.0, b));          // r = b - A x
  RC(dgx_vec_l2_norm(r, &res));
  int it = 0;
  bool converged = res <= tol;
  bool mass_front = A->kind == DGX_OP_MASS && pc.kind == 0 && A->mf->ctx->nranks >= 1 && !getenv("DGX_NO_MASS_CG_FUSION");
  while (!converged && it < max_it && std::isfinite(res)) {
    ++it;
    prev_alpha = alpha;
    dgx_vec *zz = z;
    if (pc.kind == 0) zz = r; else RC(apply_precond(pc, z, r));
    bool fused_front = false;
    if (it > 1) {
      prev_rz = r_dot_z;
      if (pc.kind == 0) r_dot_z = rr;           // z == r: r.z is the r.r the previous update already reduced
      else RC(dgx_vec_dot(r, zz, &r_dot_z));
      beta = r_dot_z / prev_rz;
    } else {
      RC(dgx_vec_dot(r, zz, &r_dot_z));
    }
    if (mass_front) {   // mass matrix, identity preconditioner: p update + M p + p.v in one kernel
      const int frc = mass_cg_front(A, p, r, beta, it == 1, v, &pv);
      if (frc == DGX_OK) fused_front = true;
      else if (frc != DGX_ERR_UNSUPPORTED) return frc;
      else mass_front = false;
    }
    if (!fused_front) {
      if (it > 1) RC(dgx_vec_sadd(p, beta, 1.0, zz));     // p = beta p + z
      else RC(dgx_vec_copy(p, zz));
      RC(dgx_op_vmult(A, v, p));
      RC(dgx_vec_dot(p, v, &pv));
    }
    alpha = r_dot_z / pv;
    RC(dgx_vec_cg_update(x, alpha, p, r, -alpha, v, &rr));   // x += alpha p ; r -= alpha v ; rr = r.r  (one pass)
    res = std::sqrt(rr);
    if (it > 1 && lz_diag) {
      lz_diag->push_back(1.0 / prev_alpha + eigen_beta_alpha);
      eigen_beta_alpha = beta / prev_alpha;
      lz_off->push_back(std::sqrt(beta) / prev_alpha);
    }
    converged = res <= tol;
  }
  if (iters) *iters = it;
  if (final_res) *final_res = res;
  if (!converged && throw_on_fail) {
    set_error("SolverCG: no convergence after " + std::to_string(it) + " steps, residual " + std::to_string(res));
    return DGX_ERR_NO_CONVERGENCE;
  }
  return DGX_OK;
}

}  // namespace

extern "C" {

// ---- MGTransferMatrixFree ----------------------------------------------------------------------------
int dgx_transfer_prolongate_and_add(dgx_vec *dst_fine, const dgx_vec *src_coarse) { return transfer(true, dst_fine, src_coarse); }
int dgx_transfer_restrict_and_add(dgx_vec *dst_coarse, const dgx_vec *src_fine) { return transfer(false, dst_coarse, src_fine); }

// SolutionTransfer::interpolate after a global refinement (refine_mesh, NS.cc:2781-2822, restricted to the case where every cell is
// refined): a DG field on the children of a cell is the embedding of the parent's polynomial -- exact, component by component.
int dgx_solution_transfer_refine(dgx_vec *dst_fine, const dgx_vec *src_coarse) {
  DGX_REQUIRE(dst_fine && src_coarse, "null argument");
  RC(dgx_vec_set(dst_fine, 0.0));
  return transfer(true, dst_fine, src_coarse);
}

// ---- PreconditionChebyshev ----------------------------------------------------------------------------
int dgx_cheb_create(dgx_op *op, int degree, double smoothing_range, int eig_cg_n_iterations, dgx_cheb **out) {
  DGX_REQUIRE(op && out, "null argument");
  DGX_REQUIRE(degree >= 1, "Chebyshev degree");
  auto *c = new dgx_cheb();
  c->op = op; c->degree = degree; c->smoothing_range = smoothing_range; c->eig_cg_n_iterations = eig_cg_n_iterations;
  for (dgx_vec **v : {&c->t, &c->x_old, &c->x_tmp}) {
    int rc = dgx_vec_create(op->mf, op->degree, op->ncomp, v);
    if (rc) { delete c; return rc; }
  }
  *out = c;
  return DGX_OK;
}

int dgx_cheb_destroy(dgx_cheb *c) {
  if (!c) return DGX_OK;
  for (dgx_vec *v : {c->t, c->x_old, c->x_tmp}) dgx_vec_destroy(v);
  delete c;
  return DGX_OK;
}

// eigenvalue estimate: eig_cg_n_iterations CG steps on D^-1 A from v = (global index mod 11) - mean
int dgx_cheb_estimate(dgx_cheb *c) {
  DGX_REQUIRE(c && c->op->inv_diag, "compute_diagonal() must precede the eigenvalue estimate");
  dgx_op *op = c->op;
  dgx_vec *w4[4];
  RC(op_work(op, w4, 4));
  dgx_vec *rhs = w4[0], *x = w4[1], *r = w4[2], *p = w4[3];
  RC(dgx_vec_set(x, 0.0));
  const long long n = rhs->n_owned, first = op->mf->first_cell * rhs->dofs_per_cell;
  cudaStream_t st = op->mf->ctx->stream;
  if (n) {
    if (rhs->dtype == DGX_F64) iota_mod11_kernel<double><<<grid_for(n), kThreads, 0, st>>>((double *)rhs->d, first, n);
    else iota_mod11_kernel<float><<<grid_for(n), kThreads, 0, st>>>((float *)rhs->d, first, n);
    count_launch();
  }
  double mean = 0;
  RC(dgx_vec_mean_value(rhs, &mean));
  if (n) {
    if (rhs->dtype == DGX_F64) add_scalar_kernel<double><<<grid_for(n), kThreads, 0, st>>>((double *)rhs->d, -mean, n);
    else add_scalar_kernel<float><<<grid_for(n), kThreads, 0, st>>>((float *)rhs->d, (float)-mean, n);
    count_launch();
  }
  std::vector<double> dg, off, ev;
  Precond pc; pc.kind = 3; pc.diag = op->inv_diag;
  int its = 0; double res = 0;
  int rc = cg_core(op, x, rhs, pc, c->eig_cg_n_iterations, 1e-10, r, p, c->t, c->x_tmp, &its, &res, &dg, &off, false);
  if (rc) return rc;
  if (dg.empty()) { c->min_eig_est = c->max_eig_est = 1.0; }
  else {
    off.resize(dg.size() - 1);
    tridiag_eigenvalues(dg, off, ev);
    c->min_eig_est = ev.front(); c->max_eig_est = ev.back();
  }
  const double max_eig = 1.2 * c->max_eig_est;
  const double alpha = c->smoothing_range > 1.0 ? max_eig / c->smoothing_range : std::min(0.9 * max_eig, c->min_eig_est);
  c->delta = (max_eig - alpha) * 0.5;
  c->theta = (max_eig + alpha) * 0.5;
  c->initialized = true;
  return DGX_OK;
}

int dgx_cheb_get_estimates(const dgx_cheb *c, double *min_eig, double *max_eig, double *theta, double *delta) {
  DGX_REQUIRE(c, "null argument");
  if (min_eig) *min_eig = c->min_eig_est;
  if (max_eig) *max_eig = c->max_eig_est;
  if (theta) *theta = c->theta;
  if (delta) *delta = c->delta;
  return DGX_OK;
}

static int cheb_iterate(dgx_cheb *c, dgx_vec *x, dgx_vec *x_old_or_null, dgx_vec *b) {
  // x holds x_1; x_old_or_null holds x_0 (null = zero).  The result ends in x.  Three buffers rotate.
  dgx_op *op = c->op;
  double rhok = c->delta / c->theta;
  const double sigma = c->theta / c->delta;
  dgx_vec *cur = x, *old = x_old_or_null;
  dgx_vec *free1 = old ? c->x_tmp : c->x_old, *free2 = old ? nullptr : c->x_tmp;
  for (int k = 0; k < c->degree - 1; ++k) {
    const double rhokp = 1.0 / (2.0 * sigma - rhok);
    const double f1 = rhokp * rhok, f2 = 2.0 * rhokp / c->delta;
    rhok = rhokp;
    dgx_vec *outv = free1;
    FusedUpdate fu;
    fu.xold = old; fu.d = op->inv_diag; fu.rhs = b; fu.a = 1.0 + f1; fu.b = -f1; fu.c = f2;
    int frc = pressure_apply_fused(op, outv, cur, fu);          // one kernel: operator + Chebyshev update
    if (frc == DGX_ERR_UNSUPPORTED) {
      RC(dgx_op_vmult(op, c->t, cur));
      RC(cheb_update(outv, cur, old, op->inv_diag, b, c->t, 1.0 + f1, -f1, f2));
    } else if (frc) return frc;
    if (old) free1 = old; else { free1 = free2; free2 = nullptr; }
    old = cur; cur = outv;
  }
  if (cur != x) RC(dgx_vec_copy(x, cur));
  return DGX_OK;
}

// PreconditionChebyshev::vmult: zero initial guess, degree - 1 operator applications
int dgx_cheb_vmult(dgx_cheb *c, dgx_vec *dst, dgx_vec *src) {
  DGX_REQUIRE(c && dst && src, "null argument");
  if (!c->initialized) RC(dgx_cheb_estimate(c));
  RC(cheb_update(dst, nullptr, nullptr, c->op->inv_diag, src, nullptr, 0.0, 0.0, 1.0 / c->theta));
  if (c->degree < 2 || std::fabs(c->delta) < 1e-40) return DGX_OK;
  return cheb_iterate(c, dst, nullptr, src);
}

// PreconditionChebyshev::step: non-zero initial guess, degree operator applications
int dgx_cheb_step(dgx_cheb *c, dgx_vec *x, dgx_vec *b) {
  DGX_REQUIRE(c && x && b, "null argument");
  if (!c->initialized) RC(dgx_cheb_estimate(c));
  RC(dgx_vec_copy(c->x_old, x));
  {
    FusedUpdate fu;
    fu.d = c->op->inv_diag; fu.rhs = b; fu.a = 1.0; fu.c = 1.0 / c->theta;
    int frc = pressure_apply_fused(c->op, x, c->x_old, fu);      // x = x_old + 1/theta D^-1 (b - A x_old)
    if (frc == DGX_ERR_UNSUPPORTED) {
      RC(dgx_op_vmult(c->op, c->t, c->x_old));
      RC(cheb_update(x, c->x_old, nullptr, c->op->inv_diag, b, c->t, 1.0, 0.0, 1.0 / c->theta));
    } else if (frc) return frc;
  }
  if (c->degree < 2 || std::fabs(c->delta) < 1e-40) return DGX_OK;
  return cheb_iterate(c, x, c->x_old, b);
}

// ---- Multigrid / PreconditionMG -------------------------------------------------------------------------
int dgx_mg_create(dgx_ctx *ctx, dgx_mesh *mesh, int degree, int level_dtype, double Re, double dt, dgx_mg **out) {
  DGX_REQUIRE(ctx && mesh && out, "null argument");
  auto *mg = new dgx_mg();
  mg->ctx = ctx; mg->mesh = mesh; mg->degree = degree; mg->Re = Re; mg->dt = dt;
  mg->n_levels = mesh->m.n_refine + 1;
  // Level distribution over the ranks.  A level is cut into equal contiguous chunks of the hierarchical order when every chunk
  // is a set of whole families of children (so that restriction / prolongation stay on the rank) and the level is large enough
  // to be worth a ghost exchange per operator application; all coarser levels are REPLICATED: every rank holds them entirely
  // and computes them redundantly (one all-gather on the way down, no communication on these levels or on the way up).
  // DGX_MG_REPLICATE_MAX = largest number of cells of a replicated level (default 32^3: an operator application on such a
  // level costs less than the latency of one exchange).
  std::vector<bool> replicated(mg->n_levels, false);
  if (ctx->nranks > 1) {
    const char *env = getenv("DGX_MG_REPLICATE_MAX");
    const long long rep_max = env ? atoll(env) : 32768;
    const long long nc = 1LL << mesh->m.dim, R = ctx->nranks;
    bool coarser_distributed = false;
    for (int l = 0; l < mg->n_levels; ++l) {
      const long long N = mesh->m.n_cells(l);
      // whole families per chunk; the coarsest distributed level only needs equal chunks of whole families towards its own children
      const bool cut_ok = N % R == 0 && (l == 0 || (N / R) % nc == 0);
      replicated[l] = !coarser_distributed && (!cut_ok || N <= rep_max);
      if (!replicated[l]) {
        if (!cut_ok) { delete mg; set_error("multigrid: level " + std::to_string(l) + " cannot be cut into whole families of children over " + std::to_string(R) + " ranks"); return DGX_ERR_UNSUPPORTED; }
        coarser_distributed = true;
      }
    }
    if (replicated[mg->n_levels - 1] && mesh->m.n_cells(mg->n_levels - 1) % R == 0) replicated[mg->n_levels - 1] = false;   // the caller's fine vectors are distributed
  }
  for (int l = 0; l < mg->n_levels; ++l) {
    dgx_mf *mf = nullptr; dgx_op *op = nullptr; dgx_vec *a = nullptr, *b = nullptr, *c = nullptr; dgx_cheb *sm = nullptr;
    int rc = mf_create(ctx, mesh, l, level_dtype, replicated[l], &mf);            // NS.cc:2353-2375
    if (!rc) rc = dgx_op_create(mf, DGX_OP_PRESSURE, degree, Re, dt, &op);        // NS.cc:2376-2379
    if (!rc) rc = dgx_vec_create(mf, degree, 1, &a);
    if (!rc) rc = dgx_vec_create(mf, degree, 1, &b);
    if (!rc) rc = dgx_vec_create(mf, degree, 1, &c);
    if (!rc && l > 0) rc = dgx_cheb_create(op, 3, 15.0, 10, &sm);                  // NS.cc:2504-2508
    if (rc) { delete mg; return rc; }
    mg->mf.push_back(mf); mg->op.push_back(op); mg->defect.push_back(a); mg->sol.push_back(b); mg->t.push_back(c); mg->smoother.push_back(sm);
  }
  int rc = dgx_vec_create(mg->mf[0], degree, 1, &mg->cg_r);
  if (!rc) rc = dgx_vec_create(mg->mf[0], degree, 1, &mg->cg_p);
  if (!rc) rc = dgx_vec_create(mg->mf[0], degree, 1, &mg->cg_v);
  if (rc) { delete mg; return rc; }
  if (cudaMalloc(&mg->d_info, (2 * dgx_mg::kLog + 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&mg->d_res, (dgx_mg::kLog + 1) * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    mg->d_info = nullptr;   // fall back to the host-sequenced coarse solve
  } else {
    mg->d_count = mg->d_info + 2 * dgx_mg::kLog;
    mg->d_tol = mg->d_res + dgx_mg::kLog;
    cudaMemsetAsync(mg->d_count, 0, sizeof(int), ctx->stream);
  }
  mg->graphs_ok = getenv("DGX_NO_MG_GRAPH") == nullptr;
  *out = mg;
  return DGX_OK;
}

int dgx_mg_destroy(dgx_mg *mg) {
  if (!mg) return DGX_OK;
  for (auto &g : mg->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (mg->d_info) cudaFree(mg->d_info);
  if (mg->d_res) cudaFree(mg->d_res);
  for (dgx_vec *v : {mg->cg_r, mg->cg_p, mg->cg_v}) dgx_vec_destroy(v);
  for (int l = 0; l < mg->n_levels; ++l) {
    dgx_cheb_destroy(mg->smoother[l]);
    dgx_vec_destroy(mg->defect[l]); dgx_vec_destroy(mg->sol[l]); dgx_vec_destroy(mg->t[l]);
    dgx_op_destroy(mg->op[l]);
    dgx_mf_destroy(mg->mf[l]);
  }
  delete mg;
  return DGX_OK;
}

int dgx_mg_n_levels(const dgx_mg *mg) { return mg ? mg->n_levels : 0; }

// smoother_data[level] of NS.cc:2501-2516 (PreconditionChebyshev::AdditionalData).  Level 0 is solved by the coarse grid solver:
// the data the reference stores there (NS.cc:2510-2512) is never used by Multigrid and is accepted and ignored here.
int dgx_mg_set_smoother(dgx_mg *mg, int level, int degree, double smoothing_range, int eig_cg_n_iterations) {
  DGX_REQUIRE(mg && level >= 0 && level < mg->n_levels, "level");
  if (level == 0) return DGX_OK;
  DGX_REQUIRE(degree >= 1 && eig_cg_n_iterations >= 0, "Chebyshev degree / eigenvalue iterations");
  dgx_cheb *c = mg->smoother[level];
  if (c->degree != degree || c->smoothing_range != smoothing_range || c->eig_cg_n_iterations != eig_cg_n_iterations) {
    c->degree = degree; c->smoothing_range = smoothing_range; c->eig_cg_n_iterations = eig_cg_n_iterations;
    c->initialized = false;
    mg->eig_cache.clear();
    ++mg->est_gen;
  }
  return DGX_OK;
}

int dgx_mg_set_dt(dgx_mg *mg, double dt) {
  DGX_REQUIRE(mg, "null argument");
  mg->dt = dt;
  for (dgx_op *op : mg->op) RC(dgx_op_set_dt(op, dt));
  return DGX_OK;
}

int dgx_mg_set_TR_BDF2_stage(dgx_mg *mg, int stage) {   // NS.cc:2941-2943, 2967-2969
  DGX_REQUIRE(mg, "null argument");
  mg->stage = stage;
  for (dgx_op *op : mg->op) RC(dgx_op_set_TR_BDF2_stage(op, stage));
  return DGX_OK;
}

// what projection_step rebuilds before every solve (NS.cc:2490-2517): level diagonals, smoother eigenvalue estimates
int dgx_mg_setup(dgx_mg *mg, double coarse_abs_tol, int coarse_max_it) {
  DGX_REQUIRE(mg, "null argument");
  if (coarse_max_it != mg->coarse_max_it) ++mg->est_gen;   // a kernel argument of the captured cycle
  ++mg->n_setups;
  mg->coarse_tol = coarse_abs_tol;
  mg->coarse_max_it = coarse_max_it;
  mg->coarse_its.clear();
  if (mg->d_info) {
    RC(mg_check_coarse_log(mg));   // also brings log_seen up to date
    mg->log_setup = mg->log_seen;
    set_scalar_kernel<<<1, 1, 0, mg->ctx->stream>>>(mg->d_tol, coarse_abs_tol);
    count_launch();
  }
  const std::vector<std::array<double, 4>> *hit = nullptr;
  if (mg->cache_estimates)
    for (auto &e : mg->eig_cache)
      if (e.first.stage == mg->stage && e.first.dt == mg->dt) hit = &e.second;
  mg->last_setup_cached = hit ? 1 : 0;
  std::vector<std::array<double, 4>> fresh(mg->n_levels);
  for (int l = 0; l < mg->n_levels; ++l) {
    RC(dgx_op_compute_diagonal(mg->op[l]));
    if (l == 0) continue;
    dgx_cheb *c = mg->smoother[l];
    if (hit) {
      c->min_eig_est = (*hit)[l][0]; c->max_eig_est = (*hit)[l][1]; c->theta = (*hit)[l][2]; c->delta = (*hit)[l][3];
      c->initialized = true;
    } else {
      c->initialized = false;
      RC(dgx_cheb_estimate(c));
      fresh[l] = {c->min_eig_est, c->max_eig_est, c->theta, c->delta};
    }
  }
  if (!hit) ++mg->est_gen;
  if (!hit && mg->cache_estimates) mg->eig_cache.push_back({{mg->stage, mg->dt}, fresh});
  return DGX_OK;
}

// 0: recompute the eigenvalue estimates on every setup, exactly as projection_step does (NS.cc:2501-2517); 1 (default): reuse them
// for an unchanged (stage, dt).  Returns through *was_cached whether the last setup hit the cache.
int dgx_mg_cache_estimates(dgx_mg *mg, int enable, int *was_cached) {
  DGX_REQUIRE(mg, "null argument");
  if (enable >= 0) { mg->cache_estimates = enable != 0; if (!enable) mg->eig_cache.clear(); }
  if (was_cached) *was_cached = mg->last_setup_cached;
  return DGX_OK;
}

static int mg_level(dgx_mg *mg, int l) {
  // input: defect[l]; output: sol[l]        (Multigrid::level_v_step)
  if (l == 0) {
    if (mg->d_info) {   // tiny level: the whole CG runs inside one CTA (no host round trips)
      int rc = pressure_coarse_cg(mg->op[0], mg->sol[0], mg->defect[0], mg->cg_r, mg->cg_p, mg->cg_v, mg->d_tol, mg->coarse_max_it, mg->d_info, mg->d_res,
                                  mg->d_count, dgx_mg::kLog);
      if (rc == DGX_OK) return DGX_OK;
      if (rc != DGX_ERR_UNSUPPORTED) return rc;
    }
    mg->graphs_ok = false;   // the host-sequenced coarse solve synchronises: the cycle cannot be captured
    RC(dgx_vec_set(mg->sol[0], 0.0));
    Precond pc;   // identity
    int its = 0; double res = 0;
    int rc = cg_core(mg->op[0], mg->sol[0], mg->defect[0], pc, mg->coarse_max_it, mg->coarse_tol, mg->cg_r, mg->cg_p, mg->cg_v, mg->t[0], &its, &res, nullptr, nullptr, true);
    mg->coarse_its.push_back(its);
    return rc;
  }
  RC(dgx_cheb_vmult(mg->smoother[l], mg->sol[l], mg->defect[l]));                  // pre-smoothing from zero
  {
    FusedUpdate fu;
    fu.rhs = mg->defect[l]; fu.c = 1.0;
    int frc = pressure_apply_fused(mg->op[l], mg->t[l], mg->sol[l], fu);           // t = defect - A sol in one kernel
    if (frc == DGX_ERR_UNSUPPORTED) {
      RC(dgx_op_vmult(mg->op[l], mg->t[l], mg->sol[l]));
      RC(dgx_vec_sadd(mg->t[l], -1.0, 1.0, mg->defect[l]));
    } else if (frc) return frc;
  }
  RC(dgx_vec_set(mg->defect[l - 1], 0.0));
  RC(dgx_transfer_restrict_and_add(mg->defect[l - 1], mg->t[l]));
  RC(mg_level(mg, l - 1));
  RC(dgx_transfer_prolongate_and_add(mg->sol[l], mg->sol[l - 1]));
  return dgx_cheb_step(mg->smoother[l], mg->sol[l], mg->defect[l]);                 // post-smoothing
}

// One V-cycle defect[L] -> sol[L].  The first application of a (stage, dt, estimates) combination runs eagerly (it allocates
// pooled work vectors and send buffers), the second one is captured into a CUDA graph, all later ones replay it: one graph
// launch instead of ~60 kernel launches, NCCL ghost exchanges and the all-gather included.  DGX_NO_MG_GRAPH=1 disables it.
int mg_cycle(dgx_mg *mg) {
  const int L = mg->n_levels - 1;
  if (!mg->graphs_ok) return mg_level(mg, L);
  dgx_mg::CycleGraph *g = nullptr;
  for (auto &e : mg->graphs)
    if (e.stage == mg->stage && e.dt == mg->dt && e.gen == mg->est_gen) g = &e;
  if (!g) {
    if (mg->graphs.size() > 16) {   // stale entries (estimates recomputed on every setup): keep the table short
      for (auto &e : mg->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
      mg->graphs.clear();
    }
    dgx_mg::CycleGraph e;
    e.stage = mg->stage; e.dt = mg->dt; e.gen = mg->est_gen; e.first_setup = mg->n_setups;
    mg->graphs.push_back(e);
    g = &mg->graphs.back();
  }
  cudaStream_t st = mg->ctx->stream;
  if (!g->exec && g->first_setup == mg->n_setups) { mg->warmed = true; return mg_level(mg, L); }   // first solve with this combination: eager
  if (!g->exec) {
    const long long l0 = g_launch_count;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); mg->graphs_ok = false; return mg_level(mg, L); }
    const int rc = mg_level(mg, L);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc || ce != cudaSuccess || !graph || !mg->graphs_ok) {   // not capturable: run eagerly from now on
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      mg->graphs_ok = false;
      if (rc) return rc;
      return mg_level(mg, L);
    }
    g->launches = g_launch_count - l0;
    g_launch_count = l0;
    const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); g->exec = nullptr; mg->graphs_ok = false; return mg_level(mg, L); }
  }
  DGX_CUDA_CHECK(cudaGraphLaunch(g->exec, st));
  count_launch((int)g->launches);
  return DGX_OK;
}

// PreconditionMG::vmult: copy_to_mg (double -> float), V-cycle, copy_from_mg
int dgx_mg_vmult(dgx_mg *mg, dgx_vec *dst, dgx_vec *src) {
  DGX_REQUIRE(mg && dst && src, "null argument");
  const int L = mg->n_levels - 1;
  RC(dgx_vec_copy(mg->defect[L], src));
  RC(mg_cycle(mg));
  return dgx_vec_copy(dst, mg->sol[L]);
}

// Reads the device-side log of the single-CTA coarse solves appended since the last check; a solve that did not converge
// surfaces here as DGX_ERR_NO_CONVERGENCE (MGCoarseGridIterativeSolver throws SolverControl::NoConvergence, NS.cc:2519-2529).
// Called at every setup and at the end of every multigrid-preconditioned dgx_solve_cg.
int mg_check_coarse_log(dgx_mg *mg) {
  if (!mg->d_info) return DGX_OK;
  int count = 0;
  DGX_CUDA_CHECK(cudaMemcpyAsync(&count, mg->d_count, sizeof(int), cudaMemcpyDeviceToHost, mg->ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(mg->ctx->stream));
  const long long seen = mg->log_seen, total = count;
  mg->log_seen = total;
  const long long fresh = std::min<long long>(total - seen, dgx_mg::kLog);
  if (fresh <= 0) return DGX_OK;
  std::vector<int> info(2 * dgx_mg::kLog);
  DGX_CUDA_CHECK(cudaMemcpy(info.data(), mg->d_info, info.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for (long long e = total - fresh; e < total; ++e) {
    const int slot = (int)(e % dgx_mg::kLog);
    if (!info[2 * slot + 1]) { set_error("coarse SolverCG: no convergence after " + std::to_string(info[2 * slot]) + " steps"); return DGX_ERR_NO_CONVERGENCE; }
  }
  return DGX_OK;
}

int dgx_mg_coarse_iterations(const dgx_mg *mg_, int *out, int capacity, int *count) {
  DGX_REQUIRE(mg_ && count, "null argument");
  dgx_mg *mg = const_cast<dgx_mg *>(mg_);
  if (mg->d_info && mg->coarse_its.empty()) {   // device-side log: the solves since the last setup
    int total = 0;
    DGX_CUDA_CHECK(cudaMemcpyAsync(&total, mg->d_count, sizeof(int), cudaMemcpyDeviceToHost, mg->ctx->stream));
    DGX_CUDA_CHECK(cudaStreamSynchronize(mg->ctx->stream));
    const long long first = std::max<long long>(mg->log_setup, (long long)total - dgx_mg::kLog);
    std::vector<int> info(2 * dgx_mg::kLog);
    DGX_CUDA_CHECK(cudaMemcpy(info.data(), mg->d_info, info.size() * sizeof(int), cudaMemcpyDeviceToHost));
    *count = (int)std::max<long long>(0, total - first);
    for (long long e = first; e < total; ++e) {
      const int slot = (int)(e % dgx_mg::kLog), i = (int)(e - first);
      if (i < capacity) out[i] = info[2 * slot];
      if (!info[2 * slot + 1]) { set_error("coarse SolverCG: no convergence after " + std::to_string(info[2 * slot]) + " steps"); return DGX_ERR_NO_CONVERGENCE; }
    }
    return DGX_OK;
  }
  *count = (int)mg->coarse_its.size();
  for (int i = 0; i < *count && i < capacity; ++i) out[i] = mg->coarse_its[i];
  return DGX_OK;
}

int dgx_mg_level_op(dgx_mg *mg, int level, dgx_op **op) {
  DGX_REQUIRE(mg && op && level >= 0 && level < mg->n_levels, "level");
  *op = mg->op[level];
  return DGX_OK;
}

// ---- Krylov solvers --------------------------------------------------------------------------------------
int dgx_solve_cg(dgx_op *A, dgx_vec *x, dgx_vec *b, int precond_kind, void *precond, int max_it, double abs_tol, int *iters, double *final_res) {
  DGX_REQUIRE(A && x && b, "null argument");
  Precond pc;
  pc.kind = precond_kind;
  if (precond_kind == 1) pc.op = precond ? (dgx_op *)precond : A;
  else if (precond_kind == 2) { pc.mg = (dgx_mg *)precond; DGX_REQUIRE(pc.mg, "multigrid handle"); }
  else DGX_REQUIRE(precond_kind == 0, "precond_kind: 0 identity, 1 Jacobi, 2 multigrid");
  dgx_vec *w[4];
  RC(op_work(A, w, 4));
  const int rc = cg_core(A, x, b, pc, max_it, abs_tol, w[0], w[1], w[2], w[3], iters, final_res, nullptr, nullptr, true);
  if (rc == DGX_OK && pc.kind == 2) return mg_check_coarse_log(pc.mg);   // a failed coarse solve must not pass silently
  return rc;
}

// SolverGMRES, deal.II >= 9.6 defaults (the reference requires 9.6.0; AdditionalData::max_basis_size = 30 IS the basis size there,
// the pre-9.6 max_n_tmp_vectors = 30 meant a basis of 28): left preconditioning, restart length max_basis, the monitored
// residual is the preconditioned one from the Givens recurrence, classical Gram-Schmidt with one re-orthogonalisation.
int dgx_solve_gmres(dgx_op *A, dgx_vec *x, dgx_vec *b, int precond_kind, void *precond, int max_it, double abs_tol, int max_basis, int *iters, double *final_res) {
  DGX_REQUIRE(A && x && b, "null argument");
  Precond pc;
  pc.kind = precond_kind;
  if (precond_kind == 1) pc.op = precond ? (dgx_op *)precond : A;
  else if (precond_kind == 2) pc.mg = (dgx_mg *)precond;
  DGX_REQUIRE(max_basis >= 1 && max_basis <= 31, "max_basis (the Gram-Schmidt kernels hold at most 32 basis vectors)");
  const int kdim = max_basis;
  std::vector<dgx_vec *> V(kdim + 1, nullptr);
  dgx_vec *w = nullptr, *tmp = nullptr;
  auto cleanup = [&]() {};   // all work vectors are pooled with the operator
#define RCG(expr) do { int rc__ = (expr); if (rc__) { cleanup(); return rc__; } } while (0)
  RCG(op_pool_vec(A, 2, &w));
  RCG(op_pool_vec(A, 3, &tmp));
  int it_total = 0;
  double rho = 0;
  bool done = false, failed = false;
  while (!done) {
    RCG(dgx_op_vmult(A, tmp, x));
    RCG(dgx_vec_sadd(tmp, -1.0, 1.0, b));
    if (!V[0]) RCG(op_pool_vec(A, 4, &V[0]));
    RCG(apply_precond(pc, V[0], tmp));
    RCG(dgx_vec_l2_norm(V[0], &rho));
    if (rho <= abs_tol) break;
    if (it_total >= max_it || !std::isfinite(rho)) { failed = true; break; }
    RCG(dgx_vec_scale(V[0], 1.0 / rho));
    std::vector<double> H((size_t)(kdim + 1) * kdim, 0.0), g(kdim + 1, 0.0), cs(kdim, 0.0), sn(kdim, 0.0);
    g[0] = rho;
    int k_used = 0;
    for (int k = 0; k < kdim; ++k) {
      const dgx_vec *jac = pc.kind == 1 ? pc.op->inv_diag : (pc.kind == 3 ? pc.diag : nullptr);
      if (A->kind == DGX_OP_VELOCITY && jac && jac->dtype == w->dtype && A->variant == 0) {
        A->post_scale = jac;                      // w = D^-1 (A v_k): the Jacobi scaling rides on the operator's store
        const int vrc = dgx_op_vmult(A, w, V[k]);
        A->post_scale = nullptr;
        RCG(vrc);
      } else {
        RCG(dgx_op_vmult(A, tmp, V[k]));
        RCG(apply_precond(pc, w, tmp));
      }
      // classical Gram-Schmidt, each sweep fused (2 passes whatever k); a second sweep only when the first one lost
      // orthogonality: norm after < 10 sqrt(eps) * norm before (Kelley's test, as in deal.II's SolverGMRES).
      // DGX_GMRES_REORTH=1 restores the unconditional second sweep.
      static const bool always_reorth = getenv("DGX_GMRES_REORTH") != nullptr;
      double hn = 0, ww = 0;
      {
        std::vector<double> h(k + 1);
        RCG(dgx_vec_orthogonalize_checked(w, V.data(), k + 1, h.data(), &ww));
        for (int j = 0; j <= k; ++j) H[(size_t)j * kdim + k] += h[j];
      }
      RCG(dgx_vec_l2_norm(w, &hn));
      if (always_reorth || hn <= 10.0 * std::sqrt(std::numeric_limits<double>::epsilon()) * std::sqrt(ww)) {
        std::vector<double> h(k + 1);
        RCG(dgx_vec_orthogonalize(w, V.data(), k + 1, h.data()));
        for (int j = 0; j <= k; ++j) H[(size_t)j * kdim + k] += h[j];
        RCG(dgx_vec_l2_norm(w, &hn));
      }
      H[(size_t)(k + 1) * kdim + k] = hn;
      for (int j = 0; j < k; ++j) {
        const double a = H[(size_t)j * kdim + k], c = H[(size_t)(j + 1) * kdim + k];
        H[(size_t)j * kdim + k] = cs[j] * a + sn[j] * c;
        H[(size_t)(j + 1) * kdim + k] = -sn[j] * a + cs[j] * c;
      }
      const double den = std::hypot(H[(size_t)k * kdim + k], H[(size_t)(k + 1) * kdim + k]);
      cs[k] = H[(size_t)k * kdim + k] / den; sn[k] = H[(size_t)(k + 1) * kdim + k] / den;
      H[(size_t)k * kdim + k] = den; H[(size_t)(k + 1) * kdim + k] = 0;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      ++it_total; k_used = k + 1;
      rho = std::fabs(g[k + 1]);
      if (rho <= abs_tol) { done = true; break; }
      if (it_total >= max_it || !std::isfinite(rho)) { failed = true; done = true; break; }
      if (hn == 0.0) break;
      if (!V[k + 1]) RCG(op_pool_vec(A, 5 + k, &V[k + 1]));
      RCG(dgx_vec_equ(V[k + 1], 1.0 / hn, w));
    }
    // back substitution and update
    std::vector<double> y(k_used, 0.0);
    for (int i = k_used - 1; i >= 0; --i) {
      double s = g[i];
      for (int j = i + 1; j < k_used; ++j) s -= H[(size_t)i * kdim + j] * y[j];
      y[i] = s / H[(size_t)i * kdim + i];
    }
    for (int j = 0; j < k_used; ++j) RCG(dgx_vec_add(x, y[j], V[j]));
  }
  cleanup();
  if (iters) *iters = it_total;
  if (final_res) *final_res = rho;
  if (failed) { set_error("SolverGMRES: no convergence after " + std::to_string(it_total) + " steps"); return DGX_ERR_NO_CONVERGENCE; }
  return DGX_OK;
}

}  // extern "C"
