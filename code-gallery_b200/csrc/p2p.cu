// Ghost import over NVLink peer memory (replaces pack kernel + grouped ncclSend/ncclRecv of vector.cu for ranks on one node).
//
// deal.II: LinearAlgebra::distributed::Vector::update_ghost_values inside MatrixFree::loop (NS.cc:1400-1414) = MPI_Isend/Irecv
// of the cells next to a rank boundary.  Here every rank owns, per MatrixFree level and cell size, a MAILBOX in its own HBM
// that its face neighbours map through CUDA IPC:
//
//     [ flags: one 64-bit sequence number per rank | slot 0: ghost cells in ghost-slot order | slot 1 ]
//
// One exchange is two kernels on every rank:
//   p2p_push_kernel    gathers the rank's boundary cells and STORES them straight into the neighbours' mailboxes over NVLink
//                      (slot = sequence number & 1); the last block to finish fences (system scope) and stores the sequence number
//                      into this rank's flag in every neighbour's mailbox.
//   p2p_unpack_kernel  waits until every neighbour's flag has reached the sequence number (acquire, system scope) and copies the
//                      slot into the ghost tail of the vector.
// No NCCL kernel, no proxy thread, no staging buffer on the sender: about 2 launch latencies + one NVLink traversal per exchange
// instead of ~25 us for a grouped send/recv; the data crosses the link once, written by the same kernel that packs it.
//
// Why two slots are enough without acknowledgements: neighbour relations are mutual and every exchange is push -> unpack in
// stream order on every rank.  Rank A's push of exchange s + 2 (which reuses the slot of s) comes after A's unpack of s + 1, which
// waited for B's push of s + 1, which B enqueued after its unpack of s: B has read slot s & 1 before A overwrites it.
// The sequence number lives in device memory and is advanced by the push kernel itself, so a captured CUDA graph (the V-cycle)
// replays exchanges correctly.
#include <algorithm>
#include <cstring>
#include <map>

#include "dgx_internal.h"

namespace dgx {

struct P2PBox {
  size_t bpc = 0;            // bytes per cell
  char *d_local = nullptr;   // this rank's mailbox
  size_t data_off = 0, slot_bytes = 0;
  int npeers = 0;
  long long n_send = 0;
  int *d_send_cells = nullptr;        // [n_send] local cell
  long long *d_send_dst = nullptr;    // [n_send] ghost slot (cell units) in the receiving rank's mailbox
  int *d_send_peer = nullptr;         // [n_send] index into d_peer_base
  char **d_peer_base = nullptr;       // [npeers] neighbours' mailboxes (IPC-mapped)
  int *d_peer_ranks = nullptr;        // [npeers]
  unsigned long long *d_seq = nullptr;
  unsigned int *d_ticket = nullptr;
  int *d_err = nullptr;
  std::vector<void *> opened;
  // asynchronous mode (copy engines): packed send buffer, host-side sequence number, one contiguous range per neighbour
  char *d_sendbuf = nullptr;
  unsigned long long h_seq = 0;
  unsigned long long *d_seqword = nullptr;   // [2]
  std::vector<char *> h_peer_base;
  std::vector<long long> send_off, send_cnt, dst_first;   // cells
  cudaEvent_t ev_packed = nullptr;
  std::vector<cudaStream_t> peer_stream;   // one per neighbour: the transfers run on separate copy engines
  std::vector<cudaEvent_t> ev_sent;
};

namespace {

constexpr size_t kFlagBytes = 1024;   // room for 128 ranks, keeps the slots 1 KiB-aligned

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename Wd>
__global__ void p2p_push_kernel(const Wd *__restrict__ v, const int *__restrict__ cells, const long long *__restrict__ dst_cell,
                                const int *__restrict__ peer, char *const *__restrict__ peer_base, size_t data_off, size_t slot_bytes,
                                long long n_send, int wpc, unsigned long long *seq, unsigned int *ticket, int npeers, int my_rank) {
  const unsigned long long s = *reinterpret_cast<volatile unsigned long long *>(seq) + 1;   // advanced by the last block only
  const size_t slot_off = data_off + (size_t)(s & 1) * slot_bytes;
  const long long total = n_send * wpc;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / wpc;
    const int w = (int)(e - i * wpc);
    Wd *dst = reinterpret_cast<Wd *>(peer_base[peer[i]] + slot_off) + dst_cell[i] * wpc + w;
    *dst = v[(long long)cells[i] * wpc + w];
  }
  // completion: every block releases its stores, the last one publishes the sequence number to the neighbours
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < npeers) st_release_sys(reinterpret_cast<unsigned long long *>(peer_base[threadIdx.x]) + my_rank, s);
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(seq) = s;
  }
}

template <typename Wd>
__global__ void p2p_unpack_kernel(Wd *__restrict__ ghost_tail, const char *__restrict__ local, size_t data_off, size_t slot_bytes,
                                  long long nwords, const unsigned long long *__restrict__ seq, const int *__restrict__ peer_ranks, int npeers,
                                  int *err) {
  const unsigned long long s = *seq;
  if ((int)threadIdx.x < npeers) {
    const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(local) + peer_ranks[threadIdx.x];
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < s) {
      if (clock64() - t0 > 6000000000LL) { atomicExch(err, 1); break; }   // ~3 s: a lost neighbour must not hang the device
    }
  }
  __syncthreads();
  const Wd *src = reinterpret_cast<const Wd *>(local + data_off + (size_t)(s & 1) * slot_bytes);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nwords; e += (long long)gridDim.x * blockDim.x) ghost_tail[e] = __ldcg(src + e);   // L2: the data came in over NVLink
}

int grid_for_words(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return (int)g;
}

void free_box(P2PBox *b) {
  if (!b) return;
  for (cudaStream_t st : b->peer_stream) cudaStreamSynchronize(st);   // outgoing transfers
  for (void *p : b->opened) cudaIpcCloseMemHandle(p);
  if (b->d_local) cudaFree(b->d_local);
  if (b->d_send_cells) cudaFree(b->d_send_cells);
  if (b->d_send_dst) cudaFree(b->d_send_dst);
  if (b->d_send_peer) cudaFree(b->d_send_peer);
  if (b->d_peer_base) cudaFree(b->d_peer_base);
  if (b->d_peer_ranks) cudaFree(b->d_peer_ranks);
  if (b->d_seq) cudaFree(b->d_seq);
  if (b->d_ticket) cudaFree(b->d_ticket);
  if (b->d_err) cudaFree(b->d_err);
  if (b->d_sendbuf) cudaFree(b->d_sendbuf);
  if (b->d_seqword) cudaFree(b->d_seqword);
  if (b->ev_packed) cudaEventDestroy(b->ev_packed);
  for (cudaEvent_t e : b->ev_sent) cudaEventDestroy(e);
  for (cudaStream_t st : b->peer_stream) cudaStreamDestroy(st);
  delete b;
}

// collective over all ranks of the context: allocate the mailbox, exchange IPC handles and ghost-slot offsets, map the neighbours
int create_box(dgx_mf *mf, size_t bpc, P2PBox **out) {
  dgx_ctx *ctx = mf->ctx;
  const int R = ctx->nranks;
  auto *b = new P2PBox();
  b->bpc = bpc;
  b->slot_bytes = ((size_t)mf->n_ghost * bpc + 1023) / 1024 * 1024;
  b->data_off = kFlagBytes;
  b->npeers = (int)mf->peers.size();
  int ok_local = (R * 8 <= (int)kFlagBytes && b->npeers <= 128) ? 1 : 0;
  if (ok_local && cudaMalloc(&b->d_local, b->data_off + 2 * b->slot_bytes) != cudaSuccess) { cudaGetLastError(); ok_local = 0; }
  if (ok_local) { cudaMemset(b->d_local, 0, b->data_off + 2 * b->slot_bytes); cudaDeviceSynchronize(); }   // zero flags before any neighbour can write them
  // piece of this rank: [ok][IPC handle][recv_first per rank (-1: not a neighbour)]
  const size_t stride = 8 + sizeof(cudaIpcMemHandle_t) + 8 * (size_t)R;
  std::vector<char> all(stride * R, 0);
  {
    char *mine = all.data() + stride * ctx->rank;
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof(h));
    if (ok_local && cudaIpcGetMemHandle(&h, b->d_local) != cudaSuccess) { cudaGetLastError(); ok_local = 0; }
    long long okv = ok_local;
    std::memcpy(mine, &okv, 8);
    std::memcpy(mine + 8, &h, sizeof(h));
    std::vector<long long> rf(R, -1);
    for (auto &p : mf->peers) rf[p.rank] = p.recv_count ? p.recv_first : -1;
    std::memcpy(mine + 8 + sizeof(h), rf.data(), 8 * (size_t)R);
  }
  char *d_all = nullptr;
  DGX_CUDA_CHECK(cudaMalloc(&d_all, all.size()));
  DGX_CUDA_CHECK(cudaMemcpyAsync(d_all, all.data(), all.size(), cudaMemcpyHostToDevice, ctx->stream));
  int rc = allgather_inplace(ctx, d_all, stride);
  if (rc) { cudaFree(d_all); free_box(b); return rc; }
  DGX_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_all);
  bool ok = true;
  for (int r = 0; r < R; ++r) { long long okv; std::memcpy(&okv, all.data() + stride * r, 8); ok = ok && okv == 1; }
  // map the neighbours; a failure here must be agreed on by all ranks, so it is checked with a second (tiny) reduction below
  std::vector<char *> peer_base;
  std::vector<int> peer_ranks;
  int map_ok = ok ? 1 : 0;
  if (ok) {
    for (auto &p : mf->peers) {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, all.data() + stride * p.rank + 8, sizeof(h));
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); map_ok = 0; break; }
      b->opened.push_back(ptr);
      peer_base.push_back((char *)ptr);
      peer_ranks.push_back(p.rank);
    }
  }
  {
    double *d = nullptr;
    DGX_CUDA_CHECK(cudaMalloc(&d, 8));
    const double v = map_ok ? 0.0 : 1.0;
    DGX_CUDA_CHECK(cudaMemcpyAsync(d, &v, 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = nccl_allreduce(ctx, d, 1, false);
    if (rc) { cudaFree(d); free_box(b); return rc; }
    double bad = 0;
    DGX_CUDA_CHECK(cudaMemcpyAsync(&bad, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    if (bad != 0.0) { free_box(b); *out = nullptr; return DGX_OK; }   // every rank falls back to NCCL
  }
  // send tables: cell, destination ghost slot on the receiver, neighbour index
  std::vector<int> cells, peer;
  std::vector<long long> dst;
  for (size_t pi = 0; pi < mf->peers.size(); ++pi) {
    const auto &p = mf->peers[pi];
    long long rf_on_peer;
    std::memcpy(&rf_on_peer, all.data() + stride * p.rank + 8 + sizeof(cudaIpcMemHandle_t) + 8 * (size_t)ctx->rank, 8);
    if (!p.send_cells.empty() && rf_on_peer < 0) { free_box(b); set_error("p2p ghost import: neighbour does not expect my cells"); return DGX_ERR_INVALID; }
    b->send_off.push_back((long long)cells.size());
    b->send_cnt.push_back((long long)p.send_cells.size());
    b->dst_first.push_back(rf_on_peer);
    for (size_t i = 0; i < p.send_cells.size(); ++i) {
      cells.push_back(p.send_cells[i]);
      dst.push_back(rf_on_peer + (long long)i);
      peer.push_back((int)pi);
    }
  }
  b->n_send = (long long)cells.size();
  b->h_peer_base = peer_base;
  auto up = [&](auto **dptr, const auto &vec) -> cudaError_t {
    using E = typename std::remove_reference<decltype(vec[0])>::type;
    if (vec.empty()) { *dptr = nullptr; return cudaSuccess; }
    cudaError_t e = cudaMalloc((void **)dptr, vec.size() * sizeof(E));
    if (e != cudaSuccess) return e;
    return cudaMemcpy((void *)*dptr, vec.data(), vec.size() * sizeof(E), cudaMemcpyHostToDevice);
  };
  DGX_CUDA_CHECK(up(&b->d_send_cells, cells));
  DGX_CUDA_CHECK(up(&b->d_send_dst, dst));
  DGX_CUDA_CHECK(up(&b->d_send_peer, peer));
  DGX_CUDA_CHECK(up(&b->d_peer_base, peer_base));
  DGX_CUDA_CHECK(up(&b->d_peer_ranks, peer_ranks));
  DGX_CUDA_CHECK(cudaMalloc(&b->d_seq, 8));
  DGX_CUDA_CHECK(cudaMemset(b->d_seq, 0, 8));
  DGX_CUDA_CHECK(cudaMalloc(&b->d_ticket, 4));
  DGX_CUDA_CHECK(cudaMemset(b->d_ticket, 0, 4));
  DGX_CUDA_CHECK(cudaMalloc(&b->d_err, 4));
  DGX_CUDA_CHECK(cudaMemset(b->d_err, 0, 4));
  DGX_CUDA_CHECK(cudaDeviceSynchronize());
  *out = b;
  return DGX_OK;
}

template <typename Wd>
int launch_exchange(dgx_vec *v, P2PBox *b, cudaStream_t stream) {
  dgx_mf *mf = v->mf;
  const int wpc = (int)(b->bpc / sizeof(Wd));
  const int threads = 256;
  if (b->n_send || b->npeers) {
    const int grid = grid_for_words(b->n_send * wpc, threads);
    p2p_push_kernel<Wd><<<grid, threads, 0, stream>>>((const Wd *)v->d, b->d_send_cells, b->d_send_dst, b->d_send_peer, b->d_peer_base, b->data_off,
                                                      b->slot_bytes, b->n_send, wpc, b->d_seq, b->d_ticket, b->npeers, mf->ctx->rank);
    count_launch();
  }
  const long long nwords = mf->n_ghost * wpc;
  {
    const int grid = grid_for_words(nwords, threads);
    Wd *tail = (Wd *)((char *)v->d + (size_t)v->n_owned * v->elem());
    p2p_unpack_kernel<Wd><<<grid, threads, 0, stream>>>(tail, b->d_local, b->data_off, b->slot_bytes, nwords, b->d_seq, b->d_peer_ranks, b->npeers, b->d_err);
    count_launch();
  }
  DGX_CUDA_CHECK(cudaGetLastError());
  return DGX_OK;
}

}  // namespace

void p2p_destroy(dgx_mf *mf) {
  for (auto &kv : mf->p2p) free_box(kv.second);
  mf->p2p.clear();
}

// 1 = done over peer memory, 0 = not available (caller uses NCCL), < 0 = error code negated
int p2p_ghost_exchange(dgx_vec *v, cudaStream_t stream, int *done) {
  *done = 0;
  dgx_mf *mf = v->mf;
  dgx_ctx *ctx = mf->ctx;
  static const bool enabled = [] { const char *e = getenv("DGX_P2P"); return !(e && e[0] == '0'); }();
  if (!enabled || ctx->p2p_off) return DGX_OK;
  const size_t bpc = (size_t)v->dofs_per_cell * v->elem();
  auto it = mf->p2p.find(bpc);
  if (it == mf->p2p.end()) {
    // first exchange of this cell size on this level: collective set-up, not inside a stream capture
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) return DGX_OK;
    P2PBox *b = nullptr;
    int rc = create_box(mf, bpc, &b);
    if (rc) return rc;
    if (!b) { ctx->p2p_off = true; return DGX_OK; }
    it = mf->p2p.emplace(bpc, b).first;
  }
  P2PBox *b = it->second;
  int rc;
  if (bpc % 16 == 0) rc = launch_exchange<uint4>(v, b, stream);
  else if (bpc % 8 == 0) rc = launch_exchange<unsigned long long>(v, b, stream);
  else rc = launch_exchange<unsigned int>(v, b, stream);
  if (rc) return rc;
  *done = 1;
  return DGX_OK;
}

namespace {
template <typename Wd>
__global__ void p2p_pack_kernel(Wd *__restrict__ out, const Wd *__restrict__ v, const int *__restrict__ cells, long long n_send, int wpc,
                                unsigned long long *seqword, unsigned long long s) {
  const long long total = n_send * wpc;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / wpc;
    out[e] = v[(long long)cells[i] * wpc + (int)(e - i * wpc)];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *seqword = s;
}
}  // namespace

// Asynchronous ghost import: the boundary cells are packed on the context's stream, the COPY ENGINES carry them into the neighbours'
// mailboxes (and then this rank's sequence number) on the communication stream while the consumer kernel already runs; it waits
// for the neighbours' sequence numbers only where it first touches a ghost cell (GhostSrc).  Own mailbox (host-side sequence
// numbers), separate from the push / unpack one.
int p2p_ghost_import_async(dgx_vec *v, GhostSrc *gs, int *done) {
  *done = 0;
  dgx_mf *mf = v->mf;
  dgx_ctx *ctx = mf->ctx;
  static const bool enabled = [] { const char *e = getenv("DGX_P2P"); const char *a = getenv("DGX_P2P_ASYNC"); return !(e && e[0] == '0') && !(a && a[0] == '0'); }();
  if (!enabled || ctx->p2p_off || ctx->nranks == 1 || mf->peers.empty() || mf->peers.size() > 32) return DGX_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(ctx->stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return DGX_OK;
  const size_t bpc = (size_t)v->dofs_per_cell * v->elem();
  const size_t key = bpc | ((size_t)1 << 62);
  auto it = mf->p2p.find(key);
  if (it == mf->p2p.end()) {
    P2PBox *b = nullptr;
    int rc = create_box(mf, bpc, &b);
    if (rc) return rc;
    if (!b) { ctx->p2p_off = true; return DGX_OK; }
    DGX_CUDA_CHECK(cudaMalloc(&b->d_sendbuf, (size_t)std::max<long long>(1, b->n_send) * bpc));
    DGX_CUDA_CHECK(cudaMalloc(&b->d_seqword, 16));
    DGX_CUDA_CHECK(cudaMemset(b->d_seqword, 0, 16));
    DGX_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_packed, cudaEventDisableTiming));
    b->peer_stream.resize(b->h_peer_base.size());
    b->ev_sent.resize(b->h_peer_base.size());
    for (size_t p = 0; p < b->h_peer_base.size(); ++p) {
      DGX_CUDA_CHECK(cudaStreamCreateWithFlags(&b->peer_stream[p], cudaStreamNonBlocking));
      DGX_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_sent[p], cudaEventDisableTiming));
    }
    it = mf->p2p.emplace(key, b).first;
  }
  P2PBox *b = it->second;
  const unsigned long long s = ++b->h_seq;
  const size_t slot_off = b->data_off + (size_t)(s & 1) * b->slot_bytes;
  cudaStream_t A = ctx->stream;
  if (s > 1)
    for (cudaEvent_t e : b->ev_sent) DGX_CUDA_CHECK(cudaStreamWaitEvent(A, e, 0));   // the previous transfers have read the send buffer
  {
    const int threads = 256;
    if (bpc % 16 == 0) {
      const int wpc = (int)(bpc / 16);
      p2p_pack_kernel<uint4><<<grid_for_words(b->n_send * wpc, threads), threads, 0, A>>>((uint4 *)b->d_sendbuf, (const uint4 *)v->d, b->d_send_cells, b->n_send, wpc, b->d_seqword + (s & 1), s);
    } else {
      const int wpc = (int)(bpc / 8);
      if (bpc % 8) return DGX_OK;
      p2p_pack_kernel<unsigned long long><<<grid_for_words(b->n_send * wpc, threads), threads, 0, A>>>((unsigned long long *)b->d_sendbuf, (const unsigned long long *)v->d, b->d_send_cells, b->n_send, wpc, b->d_seqword + (s & 1), s);
    }
    count_launch();
    DGX_CUDA_CHECK(cudaGetLastError());
  }
  DGX_CUDA_CHECK(cudaEventRecord(b->ev_packed, A));
  for (size_t p = 0; p < b->h_peer_base.size(); ++p) {
    cudaStream_t B = b->peer_stream[p];
    DGX_CUDA_CHECK(cudaStreamWaitEvent(B, b->ev_packed, 0));
    if (b->send_cnt[p])
      DGX_CUDA_CHECK(cudaMemcpyAsync(b->h_peer_base[p] + slot_off + (size_t)b->dst_first[p] * bpc, b->d_sendbuf + (size_t)b->send_off[p] * bpc,
                                     (size_t)b->send_cnt[p] * bpc, cudaMemcpyDefault, B));
    // stream order: the sequence number lands behind the data
    DGX_CUDA_CHECK(cudaMemcpyAsync(b->h_peer_base[p] + (size_t)ctx->rank * 8, b->d_seqword + (s & 1), 8, cudaMemcpyDefault, B));
    DGX_CUDA_CHECK(cudaEventRecord(b->ev_sent[p], B));
  }
  gs->base = b->d_local + slot_off;
  gs->flags = reinterpret_cast<const unsigned long long *>(b->d_local);
  gs->peer_ranks = b->d_peer_ranks;
  gs->npeers = b->npeers;
  gs->seq = s;
  *done = 1;
  return DGX_OK;
}

// ---- all-reduce of a few doubles over peer memory (dot products, norms) --------------------------------------------------------
// Replaces ncclAllReduce of 1 - 32 doubles (MPI_Allreduce in deal.II's vector dot / norm): every rank stores its values into its
// own row of a small table in EVERY rank's HBM, then its sequence number; every rank waits for all rows and adds them up in rank
// order, so all ranks get bit-identical sums.  One single-block kernel, no NCCL launch: a few microseconds instead of ~20.
struct P2PScalar {
  char *d_local = nullptr;            // [seq: R x 8 B, padded to 1 KiB][2 slots][R rows][32 doubles]
  char **d_base = nullptr;            // [R] every rank's table (own entry = d_local)
  unsigned long long *d_seq = nullptr;
  int *d_err = nullptr;
  std::vector<void *> opened;
};

namespace {
constexpr int kScalarMax = 32;

__global__ void p2p_allreduce_kernel(double *vals, int n, int is_max, char *const *__restrict__ base, unsigned long long *seq, int R, int me, int *err) {
  const unsigned long long s = *seq + 1;
  const int t = threadIdx.x;
  // my row into every rank's table
  for (int e = t; e < R * n; e += blockDim.x) {
    const int r = e / n, i = e - r * n;
    double *row = reinterpret_cast<double *>(base[r] + kFlagBytes) + ((size_t)(s & 1) * R + me) * kScalarMax;
    row[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  if (t < R) st_release_sys(reinterpret_cast<unsigned long long *>(base[t]) + me, s);
  // all rows of my table
  if (t < R) {
    const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(base[me]) + t;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < s) {
      if (clock64() - t0 > 6000000000LL) { atomicExch(err, 1); break; }
    }
  }
  __syncthreads();
  if (t < n) {
    const double *tab = reinterpret_cast<const double *>(base[me] + kFlagBytes) + (size_t)(s & 1) * R * kScalarMax;
    double a = __ldcg(tab + t);
    for (int r = 1; r < R; ++r) {
      const double b = __ldcg(tab + (size_t)r * kScalarMax + t);
      a = is_max ? fmax(a, b) : a + b;
    }
    vals[t] = a;
  }
  __syncthreads();
  if (t == 0) *seq = s;
}

int create_scalar(dgx_ctx *ctx, P2PScalar **out) {
  const int R = ctx->nranks;
  auto *b = new P2PScalar();
  const size_t bytes = kFlagBytes + 2 * (size_t)R * kScalarMax * 8;
  int ok_local = R * 8 <= (int)kFlagBytes ? 1 : 0;
  if (ok_local && cudaMalloc(&b->d_local, bytes) != cudaSuccess) { cudaGetLastError(); ok_local = 0; }
  if (ok_local) { cudaMemset(b->d_local, 0, bytes); cudaDeviceSynchronize(); }
  const size_t stride = 8 + sizeof(cudaIpcMemHandle_t);
  std::vector<char> all(stride * R, 0);
  {
    cudaIpcMemHandle_t h;
    std::memset(&h, 0, sizeof(h));
    if (ok_local && cudaIpcGetMemHandle(&h, b->d_local) != cudaSuccess) { cudaGetLastError(); ok_local = 0; }
    long long okv = ok_local;
    std::memcpy(all.data() + stride * ctx->rank, &okv, 8);
    std::memcpy(all.data() + stride * ctx->rank + 8, &h, sizeof(h));
  }
  char *d_all = nullptr;
  DGX_CUDA_CHECK(cudaMalloc(&d_all, all.size()));
  DGX_CUDA_CHECK(cudaMemcpyAsync(d_all, all.data(), all.size(), cudaMemcpyHostToDevice, ctx->stream));
  int rc = allgather_inplace(ctx, d_all, stride);
  if (rc) { cudaFree(d_all); return rc; }
  DGX_CUDA_CHECK(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_all);
  int map_ok = 1;
  for (int r = 0; r < R; ++r) { long long okv; std::memcpy(&okv, all.data() + stride * r, 8); if (okv != 1) map_ok = 0; }
  std::vector<char *> base(R, nullptr);
  for (int r = 0; r < R && map_ok; ++r) {
    if (r == ctx->rank) { base[r] = b->d_local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all.data() + stride * r + 8, sizeof(h));
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); map_ok = 0; break; }
    b->opened.push_back(ptr);
    base[r] = (char *)ptr;
  }
  {   // agree on the outcome with NCCL (the last time it is used for a scalar)
    double *d = nullptr;   // not the context's scratch: the caller's partial sums live there
    DGX_CUDA_CHECK(cudaMalloc(&d, 8));
    const double v = map_ok ? 0.0 : 1.0;
    DGX_CUDA_CHECK(cudaMemcpyAsync(d, &v, 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = nccl_allreduce(ctx, d, 1, false);
    if (rc) { cudaFree(d); return rc; }
    double bad = 0;
    DGX_CUDA_CHECK(cudaMemcpyAsync(&bad, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    DGX_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    if (bad != 0.0) {
      for (void *p : b->opened) cudaIpcCloseMemHandle(p);
      if (b->d_local) cudaFree(b->d_local);
      delete b;
      *out = nullptr;
      return DGX_OK;
    }
  }
  DGX_CUDA_CHECK(cudaMalloc(&b->d_base, R * sizeof(char *)));
  DGX_CUDA_CHECK(cudaMemcpy(b->d_base, base.data(), R * sizeof(char *), cudaMemcpyHostToDevice));
  DGX_CUDA_CHECK(cudaMalloc(&b->d_seq, 8));
  DGX_CUDA_CHECK(cudaMemset(b->d_seq, 0, 8));
  DGX_CUDA_CHECK(cudaMalloc(&b->d_err, 4));
  DGX_CUDA_CHECK(cudaMemset(b->d_err, 0, 4));
  DGX_CUDA_CHECK(cudaDeviceSynchronize());
  *out = b;
  return DGX_OK;
}
}  // namespace

void p2p_scalar_destroy(dgx_ctx *ctx) {
  P2PScalar *b = ctx->p2ps;
  if (!b) return;
  for (void *p : b->opened) cudaIpcCloseMemHandle(p);
  if (b->d_local) cudaFree(b->d_local);
  if (b->d_base) cudaFree(b->d_base);
  if (b->d_seq) cudaFree(b->d_seq);
  if (b->d_err) cudaFree(b->d_err);
  delete b;
  ctx->p2ps = nullptr;
}

// *done = 1 when the reduction was enqueued on the context's stream, 0 when the caller must use NCCL
int p2p_allreduce(dgx_ctx *ctx, double *dvals, int n, bool is_max, int *done) {
  *done = 0;
  static const bool enabled = [] { const char *e = getenv("DGX_P2P"); return !(e && e[0] == '0'); }();
  if (!enabled || ctx->p2p_off || n > kScalarMax || ctx->nranks > 64) return DGX_OK;
  if (!ctx->p2ps) {
    if (ctx->p2ps_tried) return DGX_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) return DGX_OK;
    ctx->p2ps_tried = true;
    P2PScalar *b = nullptr;
    int rc = create_scalar(ctx, &b);
    if (rc) return rc;
    if (!b) return DGX_OK;
    ctx->p2ps = b;
  }
  P2PScalar *b = ctx->p2ps;
  p2p_allreduce_kernel<<<1, 128, 0, ctx->stream>>>(dvals, n, is_max ? 1 : 0, b->d_base, b->d_seq, ctx->nranks, ctx->rank, b->d_err);
  count_launch();
  DGX_CUDA_CHECK(cudaGetLastError());
  *done = 1;
  return DGX_OK;
}

}  // namespace dgx
