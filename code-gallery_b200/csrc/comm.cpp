// NCCL plumbing, resolved at run time with dlopen so that libdgx.so loads on machines / processes
// without NCCL and shares the copy torch has already loaded.  Replaces the MPI calls deal.II makes
// for the reference: Partitioner Isend/Irecv ghost import -> grouped
// ncclSend/ncclRecv; MPI_Allreduce of dot products / norms -> ncclAllReduce of a few doubles.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "dgx_internal.h"

namespace dgx {
namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) return;
#define LOAD(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym))
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(AllGather, "ncclAllGather");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.AllReduce && a.AllGather && a.GroupStart && a.GroupEnd;
  });
  return a;
}

#define DGX_NCCL_CHECK(expr)                                                                  \
  do {                                                                                        \
    ncclResult_t r__ = (expr);                                                                \
    if (r__ != ncclSuccess) {                                                                 \
      set_error(std::string(#expr) + ": " + (api().GetErrorString ? api().GetErrorString(r__) : "nccl error")); \
      return DGX_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)
}  // namespace

int comm_unique_id(void *out) {
  if (!api().ok) { set_error("NCCL library not found (dlopen libnccl.so.2)"); return DGX_ERR_NCCL; }
  static_assert(sizeof(ncclUniqueId) <= DGX_NCCL_ID_BYTES, "id size");
  ncclUniqueId id;
  DGX_NCCL_CHECK(api().GetUniqueId(&id));
  std::memset(out, 0, DGX_NCCL_ID_BYTES);
  std::memcpy(out, &id, sizeof(id));
  return DGX_OK;
}

int comm_init(dgx_ctx *ctx, const void *idbytes) {
  if (ctx->nranks == 1) return DGX_OK;
  if (!api().ok) { set_error("NCCL library not found (dlopen libnccl.so.2)"); return DGX_ERR_NCCL; }
  DGX_REQUIRE(idbytes, "nccl_id required for nranks > 1");
  ncclUniqueId id;
  std::memcpy(&id, idbytes, sizeof(id));
  ncclComm_t comm;
  DGX_NCCL_CHECK(api().CommInitRank(&comm, ctx->nranks, id, ctx->rank));
  ctx->nccl_comm = comm;
  return DGX_OK;
}

int comm_destroy(dgx_ctx *ctx) {
  p2p_scalar_destroy(ctx);
  if (ctx->nccl_comm) api().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  return DGX_OK;
}

int nccl_allreduce(dgx_ctx *ctx, double *dvals, int n, bool is_max) {
  if (ctx->nranks == 1) return DGX_OK;
  DGX_NCCL_CHECK(api().AllReduce(dvals, dvals, n, ncclFloat64, is_max ? ncclMax : ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return DGX_OK;
}

// dot products / norms: over peer memory when the ranks can map each other (p2p.cu), else NCCL
static int allreduce_any(dgx_ctx *ctx, double *dvals, int n, bool is_max) {
  if (ctx->nranks == 1) return DGX_OK;
  int done = 0;
  int rc = p2p_allreduce(ctx, dvals, n, is_max, &done);
  if (rc) return rc;
  if (done) return DGX_OK;
  return nccl_allreduce(ctx, dvals, n, is_max);
}
int allreduce_sum(dgx_ctx *ctx, double *dvals, int n) { return allreduce_any(ctx, dvals, n, false); }
int allreduce_max(dgx_ctx *ctx, double *dvals, int n) { return allreduce_any(ctx, dvals, n, true); }

int allgather_inplace(dgx_ctx *ctx, void *buf, size_t bytes_per_rank) {
  if (ctx->nranks == 1 || bytes_per_rank == 0) return DGX_OK;
  DGX_NCCL_CHECK(api().AllGather((const char *)buf + (size_t)ctx->rank * bytes_per_rank, buf, bytes_per_rank, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return DGX_OK;
}

// grouped point-to-point: send[i] -> peers, recv[i] <- peers (bytes)
int comm_exchange(dgx_ctx *ctx, cudaStream_t stream, int npeers, const int *ranks, const void *const *sendbuf, const size_t *sendbytes,
                  void *const *recvbuf, const size_t *recvbytes) {
  if (ctx->nranks == 1 || npeers == 0) return DGX_OK;
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  DGX_NCCL_CHECK(api().GroupStart());
  ncclResult_t bad = ncclSuccess;   // never return between GroupStart and GroupEnd: the group would stay open
  for (int i = 0; i < npeers && bad == ncclSuccess; ++i) {
    if (sendbytes[i]) bad = api().Send(sendbuf[i], sendbytes[i], ncclChar, ranks[i], comm, stream);
    if (bad == ncclSuccess && recvbytes[i]) bad = api().Recv(recvbuf[i], recvbytes[i], ncclChar, ranks[i], comm, stream);
  }
  const ncclResult_t end = api().GroupEnd();
  DGX_NCCL_CHECK(bad);
  DGX_NCCL_CHECK(end);
  return DGX_OK;
}

}  // namespace dgx
