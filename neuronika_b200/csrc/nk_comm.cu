// NCCL behind the C ABI (SURVEY.md 8-b "nk_allreduce_sum(ctx, ptr, n, dtype)", ctx owns the communicator; 8-e: the
// gradient buffers of the replicas are summed after backward).  A host in any language drives data parallel through
// these four calls -- no Python, no torch:
//     rank 0: nk_comm_unique_id(id)  -> ship the 128 bytes to the other ranks (file, socket, MPI ...)
//     all   : nk_comm_init_rank(ctx, world, rank, id);  ...backward...;  nk_allreduce_sum(ctx, grads, n, NK_F32)
// libnccl is bound at run time (dlopen "libnccl.so.2"): the library keeps loading -- and every other entry point keeps
// working -- on a host without NCCL, and inside a process that already carries a NCCL (e.g. torch's bundled copy) the
// same instance is used.  Collectives are enqueued on the context stream, ordered with the kernels around them.
#include <dlfcn.h>
#include <string.h>

#include "nk_internal.cuh"

namespace {

// the handful of NCCL declarations used here (nccl.h 2.27: values are ABI-stable across 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSumOp = 0 };
enum { ncclFloat32T = 7, ncclBfloat16T = 9 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool tried = false;
};
NcclApi g_nccl;

const char* load_nccl() {  // returns NULL on success, else what failed
  if (g_nccl.handle) return nullptr;
  if (g_nccl.tried) return "libnccl.so.2 could not be loaded";
  g_nccl.tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return "libnccl.so.2 could not be loaded";
#define NK_SYM(field, name)                                             \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (!g_nccl.field) return "libnccl lacks " name
  NK_SYM(GetUniqueId, "ncclGetUniqueId");
  NK_SYM(CommInitRank, "ncclCommInitRank");
  NK_SYM(CommDestroy, "ncclCommDestroy");
  NK_SYM(AllReduce, "ncclAllReduce");
  NK_SYM(GetErrorString, "ncclGetErrorString");
  NK_SYM(GetVersion, "ncclGetVersion");
#undef NK_SYM
  g_nccl.handle = h;
  return nullptr;
}

int nccl_fail(nk_ctx* ctx, const char* what, int rc) {
  return nk_set_error(ctx, NK_ERR_NCCL, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
}

}  // namespace

extern "C" {

int nk_comm_unique_id(nk_ctx* ctx, void* id128) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, id128 != nullptr, "nk_comm_unique_id: NULL buffer");
  if (const char* e = load_nccl()) return nk_set_error(ctx, NK_ERR_NCCL, "nk_comm_unique_id: %s", e);
  ncclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != ncclSuccess) return nccl_fail(ctx, "ncclGetUniqueId", rc);
  memcpy(id128, id.internal, 128);
  return NK_OK;
}

int nk_comm_init_rank(nk_ctx* ctx, int world, int rank, const void* id128) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world && id128, "nk_comm_init_rank: bad world %d / rank %d", world, rank);
  NK_REQUIRE(ctx, ctx->comm == nullptr, "nk_comm_init_rank: the context already owns a communicator");
  if (const char* e = load_nccl()) return nk_set_error(ctx, NK_ERR_NCCL, "nk_comm_init_rank: %s", e);
  NK_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  int rc = g_nccl.CommInitRank(&comm, world, id, rank);
  if (rc != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", rc);
  ctx->comm = comm;
  ctx->comm_world = world;
  ctx->comm_rank = rank;
  return NK_OK;
}

int nk_comm_destroy(nk_ctx* ctx) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  if (ctx->comm && g_nccl.CommDestroy) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  }
  ctx->comm = nullptr;
  ctx->comm_world = 0;
  return NK_OK;
}

int nk_comm_world(nk_ctx* ctx) { return ctx ? (ctx->comm ? ctx->comm_world : 0) : 0; }
int nk_comm_rank(nk_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }

int nk_allreduce_sum(nk_ctx* ctx, void* ptr, size_t n, int dtype) {
  if (!ctx) return NK_ERR_INVALID_ARG;
  NK_REQUIRE(ctx, nk_dtype_ok(dtype), "nk_allreduce_sum: bad dtype %d", dtype);
  if (!ctx->comm) return nk_set_error(ctx, NK_ERR_NCCL, "nk_allreduce_sum: no communicator (call nk_comm_init_rank first)");
  if (n == 0) return NK_OK;
  NK_REQUIRE(ctx, ptr != nullptr, "nk_allreduce_sum: NULL pointer");
  int rc = g_nccl.AllReduce(ptr, ptr, n, dtype == NK_BF16 ? ncclBfloat16T : ncclFloat32T, ncclSumOp,
                            static_cast<ncclComm_t>(ctx->comm), ctx->stream);
  if (rc != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce", rc);
  return NK_OK;
}

}  // extern "C"
