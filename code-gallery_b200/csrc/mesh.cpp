// Cartesian refine_global hierarchy in deal.II cell order + MatrixFree-style setup tables.
//
// Replaces what the reference gets from Triangulation::refine_global / DoFHandler::distribute_dofs /
// MatrixFree::reinit (NS.cc:2261-2270, 2282-2283, 2329, 2348-2375) for uniformly refined Cartesian
// meshes: coarse cells lexicographic, Morton order below each coarse cell, DG dofs
// numbered cell by cell.  The "MatrixFree" here is SoA: one int32 face-neighbour column per face,
// Cartesian metric as 3 doubles, ghost cells appended after the owned range.
#include <algorithm>
#include <map>
#include <set>

#include "dgx_internal.h"

namespace dgx {

long long Mesh::n_cells(int level) const {
  long long n = 1;
  for (int d = 0; d < dim; ++d) n *= (long long)cells_dir(level, d);
  return n;
}

static inline long long morton_encode(const int *c, int nbits, int dim) {
  long long out = 0;
  for (int b = 0; b < nbits; ++b)
    for (int d = 0; d < dim; ++d) out |= (long long)((c[d] >> b) & 1) << (dim * b + d);
  return out;
}

void Mesh::coords_of(int level, long long idx, int *c) const {
  const long long per = 1LL << (dim * level);
  long long cidx = idx / per, loc = idx % per;
  for (int d = 0; d < dim; ++d) {
    int l = 0;
    for (int b = 0; b < level; ++b) l |= (int)((loc >> (dim * b + d)) & 1) << b;
    c[d] = (int)((cidx % cells0[d]) << level) + l;
    cidx /= cells0[d];
  }
}

long long Mesh::index_of(int level, const int *c) const {
  int loc[MAX_DIM];
  long long cidx = 0, stride = 1;
  const int mask = (1 << level) - 1;
  for (int d = 0; d < dim; ++d) {
    cidx += (long long)(c[d] >> level) * stride;
    stride *= cells0[d];
    loc[d] = c[d] & mask;
  }
  return cidx * (1LL << (dim * level)) + morton_encode(loc, level, dim);
}

long long Mesh::neighbor(int level, long long idx, int f) const {
  int c[MAX_DIM];
  coords_of(level, idx, c);
  const int d = f / 2, s = f % 2, nd = cells_dir(level, d);
  c[d] += 2 * s - 1;
  if (c[d] < 0 || c[d] >= nd) {
    if (!periodic[d]) return -1 - (long long)bids[f];
    c[d] = (c[d] + nd) % nd;
  }
  return index_of(level, c);
}

}  // namespace dgx

using namespace dgx;

extern "C" {

int dgx_mesh_create_cartesian(int dim, const int *cells0, int n_refine, const double *lo, const double *hi,
                              int periodic_mask, const int *boundary_ids, dgx_mesh **out) {
  DGX_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  DGX_REQUIRE(out && cells0 && lo && hi, "null argument");
  DGX_REQUIRE(n_refine >= 0 && n_refine <= 12, "n_refine out of range");
  auto *m = new dgx_mesh();
  m->m.dim = dim;
  m->m.n_refine = n_refine;
  for (int d = 0; d < dim; ++d) {
    if (cells0[d] < 1 || !(hi[d] > lo[d])) { delete m; set_error("bad coarse grid / box"); return DGX_ERR_INVALID; }
    m->m.cells0[d] = cells0[d];
    m->m.lo[d] = lo[d];
    m->m.hi[d] = hi[d];
    m->m.periodic[d] = (periodic_mask >> d) & 1;
  }
  for (int f = 0; f < 2 * dim; ++f) m->m.bids[f] = boundary_ids ? boundary_ids[f] : f;
  *out = m;
  return DGX_OK;
}

int dgx_mesh_destroy(dgx_mesh *mesh) { delete mesh; return DGX_OK; }

int dgx_mesh_set_vertices(dgx_mesh *mesh, int level, const double *coords, long long n_values) {
  DGX_REQUIRE(mesh && coords, "null argument");
  Mesh &m = mesh->m;
  if (level < 0) level = m.n_refine;
  DGX_REQUIRE(level <= m.n_refine, "level");
  long long nv = 1;
  for (int d = 0; d < m.dim; ++d) nv *= m.cells_dir(level, d) + 1;
  DGX_REQUIRE(n_values == nv * m.dim, "coords must hold (cells + 1)^dim vertices of dim doubles each");
  if ((int)m.vertices.size() <= m.n_refine) m.vertices.resize(m.n_refine + 1);
  m.vertices[level].assign(coords, coords + n_values);
  return DGX_OK;
}

int dgx_mesh_is_general(const dgx_mesh *mesh, int level) {
  if (!mesh) return 0;
  return mesh->m.general(level < 0 ? mesh->m.n_refine : level) ? 1 : 0;
}

int dgx_mesh_n_levels(const dgx_mesh *mesh) { return mesh ? mesh->m.n_refine + 1 : 0; }

long long dgx_mesh_n_cells(const dgx_mesh *mesh, int level) {
  if (!mesh) return 0;
  if (level < 0) level = mesh->m.n_refine;
  return mesh->m.n_cells(level);
}

int dgx_mesh_cell_coords(const dgx_mesh *mesh, int level, long long first, long long count, int *out) {
  DGX_REQUIRE(mesh && out, "null argument");
  if (level < 0) level = mesh->m.n_refine;
  DGX_REQUIRE(first >= 0 && first + count <= mesh->m.n_cells(level), "cell range");
  for (long long i = 0; i < count; ++i) mesh->m.coords_of(level, first + i, out + i * mesh->m.dim);
  return DGX_OK;
}

int dgx_mesh_neighbors(const dgx_mesh *mesh, int level, long long first, long long count, long long *out) {
  DGX_REQUIRE(mesh && out, "null argument");
  if (level < 0) level = mesh->m.n_refine;
  DGX_REQUIRE(first >= 0 && first + count <= mesh->m.n_cells(level), "cell range");
  const int nf = 2 * mesh->m.dim;
  for (long long i = 0; i < count; ++i)
    for (int f = 0; f < nf; ++f) out[i * nf + f] = mesh->m.neighbor(level, first + i, f);
  return DGX_OK;
}

// ---- MatrixFree-style setup ------------------------------------------------------------------
int dgx_mf_create(dgx_ctx *ctx, dgx_mesh *mesh, int level, int dtype, dgx_mf **out) { return dgx::mf_create(ctx, mesh, level, dtype, false, out); }

}  // extern "C"

int dgx::mf_create(dgx_ctx *ctx, dgx_mesh *mesh, int level, int dtype, bool replicated, dgx_mf **out) {
  DGX_REQUIRE(ctx && mesh && out, "null argument");
  DGX_REQUIRE(dtype == DGX_F64 || dtype == DGX_F32, "dtype");
  if (level < 0) level = mesh->m.n_refine;
  DGX_REQUIRE(level <= mesh->m.n_refine, "level");
  const Mesh &m = mesh->m;
  auto mf = std::make_unique<dgx_mf>();
  mf->ctx = ctx; mf->mesh = mesh; mf->level = level; mf->dtype = dtype; mf->dim = m.dim;
  const long long N = m.n_cells(level);
  DGX_REQUIRE(N < (1LL << 31), "too many cells");
  mf->replicated = replicated && ctx->nranks > 1;
  const int R = mf->replicated ? 1 : ctx->nranks, r = mf->replicated ? 0 : ctx->rank;
  auto bound = [&](int k) { return (N * k) / R; };
  mf->n_global = N;
  mf->first_cell = bound(r);
  mf->n_owned = bound(r + 1) - bound(r);
  for (int d = 0; d < m.dim; ++d) mf->h[d] = m.h(level, d);
  mf->general = m.general(level);
  const int nf = 2 * m.dim;
  const long long no = mf->n_owned;
  auto owner_of = [&](long long g) {
    int k = (int)((g * R) / N);
    while (k + 1 < R && bound(k + 1) <= g) ++k;
    while (k > 0 && bound(k) > g) --k;
    return k;
  };
  // global neighbours of owned cells
  std::vector<long long> gn((size_t)no * nf);
  bool all_interior = true;
#pragma omp parallel for schedule(static) reduction(&& : all_interior)
  for (long long c = 0; c < no; ++c)
    for (int f = 0; f < nf; ++f) {
      long long g = m.neighbor(level, mf->first_cell + c, f);
      gn[(size_t)c * nf + f] = g;
      if (g < 0) all_interior = false;
    }
  mf->all_interior = all_interior;
  // bounding box of the owned cells; the marching kernels need owned set == box
  {
    int lo[MAX_DIM] = {1 << 30, 1 << 30, 1 << 30}, hi[MAX_DIM] = {-1, -1, -1};
    for (long long c = 0; c < no; ++c) {
      int cc[MAX_DIM] = {0, 0, 0};
      m.coords_of(level, mf->first_cell + c, cc);
      for (int d = 0; d < m.dim; ++d) { lo[d] = std::min(lo[d], cc[d]); hi[d] = std::max(hi[d], cc[d]); }
    }
    long long vol = no > 0 ? 1 : 0;
    for (int d = 0; d < m.dim; ++d) { mf->box_lo[d] = no > 0 ? lo[d] : 0; mf->box_ext[d] = no > 0 ? hi[d] - lo[d] + 1 : 0; vol *= mf->box_ext[d]; }
    mf->box_ok = no > 0 && vol == no;
  }
  // ghost set (sorted by global index => grouped by owner rank)
  std::vector<long long> ghosts;
  if (R > 1) {
    for (size_t i = 0; i < gn.size(); ++i) {
      long long g = gn[i];
      if (g >= 0 && (g < mf->first_cell || g >= mf->first_cell + no)) ghosts.push_back(g);
    }
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  }
  mf->ghost_global = ghosts;
  mf->n_ghost = (long long)ghosts.size();
  mf->nbr_host.assign((size_t)nf * no, 0);
  for (long long c = 0; c < no; ++c)
    for (int f = 0; f < nf; ++f) {
      long long g = gn[(size_t)c * nf + f];
      int loc;
      if (g < 0) loc = (int)g;
      else if (g >= mf->first_cell && g < mf->first_cell + no) loc = (int)(g - mf->first_cell);
      else loc = (int)(no + (std::lower_bound(ghosts.begin(), ghosts.end(), g) - ghosts.begin()));
      mf->nbr_host[(size_t)f * no + c] = loc;
    }
  // exchange plan.  Receive side: ghost slots grouped by owner.  Send side: my cells adjacent to a
  // cell of the peer, sorted by global index (the peer derives the same list as its ghost range).
  if (R > 1) {
    std::map<int, std::set<long long>> send;
    for (long long c = 0; c < no; ++c)
      for (int f = 0; f < nf; ++f) {
        long long g = gn[(size_t)c * nf + f];
        if (g >= 0 && (g < mf->first_cell || g >= mf->first_cell + no)) send[owner_of(g)].insert(mf->first_cell + c);
      }
    std::map<int, std::pair<long long, long long>> recv;
    for (size_t i = 0; i < ghosts.size(); ++i) {
      int o = owner_of(ghosts[i]);
      auto it = recv.find(o);
      if (it == recv.end()) recv[o] = std::make_pair((long long)i, 1LL);
      else it->second.second++;
    }
    std::set<int> peer_ranks;
    for (auto &kv : send) peer_ranks.insert(kv.first);
    for (auto &kv : recv) peer_ranks.insert(kv.first);
    for (int p : peer_ranks) {
      dgx_mf::Peer peer;
      peer.rank = p;
      if (send.count(p)) for (long long g : send[p]) peer.send_cells.push_back((int)(g - mf->first_cell));
      if (recv.count(p)) { peer.recv_first = recv[p].first; peer.recv_count = recv[p].second; }
      mf->peers.push_back(std::move(peer));
    }
  }
  // device copies (skipped when there is no device: host-only table queries still work)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
    DGX_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!mf->nbr_host.empty()) {
      DGX_CUDA_CHECK(cudaMalloc(&mf->d_nbr, mf->nbr_host.size() * sizeof(int)));
      DGX_CUDA_CHECK(cudaMemcpy(mf->d_nbr, mf->nbr_host.data(), mf->nbr_host.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    for (auto &p : mf->peers)
      if (!p.send_cells.empty()) {
        DGX_CUDA_CHECK(cudaMalloc(&p.d_send_cells, p.send_cells.size() * sizeof(int)));
        DGX_CUDA_CHECK(cudaMemcpy(p.d_send_cells, p.send_cells.data(), p.send_cells.size() * sizeof(int), cudaMemcpyHostToDevice));
      }
  } else {
    cudaGetLastError();
  }
  *out = mf.release();
  return DGX_OK;
}

extern "C" {

int dgx_mf_destroy(dgx_mf *mf) {
  if (!mf) return DGX_OK;
  if (mf->d_nbr) cudaFree(mf->d_nbr);
  if (mf->d_jobs) cudaFree(mf->d_jobs);
  if (mf->d_jobs_async) cudaFree(mf->d_jobs_async);
  dgx::p2p_destroy(mf);
  for (auto &kv : mf->bfaces) if (kv.second.d) cudaFree(kv.second.d);
  dgx::general_metric_destroy(mf);
  if (mf->d_cellX) cudaFree(mf->d_cellX);
  for (auto &p : mf->peers) if (p.d_send_cells) cudaFree(p.d_send_cells);
  delete mf;
  return DGX_OK;
}

long long dgx_mf_n_owned_cells(const dgx_mf *mf) { return mf ? mf->n_owned : 0; }
long long dgx_mf_n_ghost_cells(const dgx_mf *mf) { return mf ? mf->n_ghost : 0; }
long long dgx_mf_first_owned_cell(const dgx_mf *mf) { return mf ? mf->first_cell : 0; }

int dgx_mf_local_neighbors(const dgx_mf *mf, int *out) {
  DGX_REQUIRE(mf && out, "null argument");
  std::copy(mf->nbr_host.begin(), mf->nbr_host.end(), out);
  return DGX_OK;
}

int dgx_mf_n_peers(const dgx_mf *mf) { return mf ? (int)mf->peers.size() : 0; }

int dgx_mf_peer_info(const dgx_mf *mf, int i, int *rank, long long *n_send, long long *recv_first, long long *recv_count) {
  DGX_REQUIRE(mf && i >= 0 && i < (int)mf->peers.size(), "peer index");
  const auto &p = mf->peers[i];
  if (rank) *rank = p.rank;
  if (n_send) *n_send = (long long)p.send_cells.size();
  if (recv_first) *recv_first = p.recv_first;
  if (recv_count) *recv_count = p.recv_count;
  return DGX_OK;
}

int dgx_mf_peer_send_cells(const dgx_mf *mf, int i, int *out) {
  DGX_REQUIRE(mf && out && i >= 0 && i < (int)mf->peers.size(), "peer index");
  std::copy(mf->peers[i].send_cells.begin(), mf->peers[i].send_cells.end(), out);
  return DGX_OK;
}

int dgx_mf_ghost_cells(const dgx_mf *mf, long long *out) {
  DGX_REQUIRE(mf && out, "null argument");
  std::copy(mf->ghost_global.begin(), mf->ghost_global.end(), out);
  return DGX_OK;
}

}  // extern "C"
