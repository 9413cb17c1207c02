// rsc_comm.cu -- the all-reduce of the sharded path inside the library: NCCL over NVLink/NVSwitch.
//
// The path shards by point range: per-candidate counts, the gathered coordinates of
// the minimal sets, K5 hit counts and the refit's inlier-mask words are int32 buffers summed over the
// ranks -- small messages (16 KB ... a few MB), so the collective is latency bound and one
// ncclAllReduce on the context stream per exchange is the right tool (no host round trip, ordered
// with the kernels around it).  libnccl.so.2 is dlopen-ed on first use: the library loads and every
// single-GPU entry point works on a machine without NCCL, and a process that already mapped an NCCL
// (torch's bundled one) shares it instead of loading a second copy.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "rsc_common.cuh"

namespace rsc {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.h || !api.err.empty()) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names)
    if ((api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!api.h) {
    api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return &api;
  }
  auto sym = [&](const char* s) {
    void* p = dlsym(api.h, s);
    if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + s;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return &api;
}

// the rsc_allreduce_fn the library installs for itself once a communicator exists
static int32_t nccl_allreduce(void* user, void* d_buf, int64_t count, void* stream) {
  rsc_ctx* ctx = static_cast<rsc_ctx*>(user);
  NcclApi* a = nccl_api();
  if (!ctx->comm || !a->AllReduce) return 1;
  const ncclResult_t r = a->AllReduce(d_buf, d_buf, (size_t)count, ncclInt32, ncclSum, static_cast<ncclComm_t>(ctx->comm),
                                      static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) {
    ctx->err = std::string("ncclAllReduce: ") + (a->GetErrorString ? a->GetErrorString(r) : "error");
    return 1;
  }
  ++ctx->allreduce_calls;
  ctx->allreduce_bytes += count * 4;
  return 0;
}

}  // namespace rsc

using namespace rsc;

extern "C" {

int32_t rsc_comm_unique_id(void* out) {
  if (!out) return RSC_E_ARG;
  NcclApi* a = nccl_api();
  if (!a->GetUniqueId) return RSC_E_NCCL;
  static_assert(sizeof(ncclUniqueId) == RSC_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return RSC_E_NCCL;
  memcpy(out, &id, sizeof(id));
  return RSC_OK;
}

int32_t rsc_ctx_comm_init(rsc_ctx* ctx, const void* unique_id, int32_t rank, int32_t nranks) {
  if (!ctx) return RSC_E_ARG;
  if (!unique_id || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, RSC_E_ARG, "comm_init: bad rank / nranks / id");
  if (ctx->comm) return fail(ctx, RSC_E_STATE, "comm_init: the context already has a communicator");
  NcclApi* a = nccl_api();
  if (!a->CommInitRank) return fail(ctx, RSC_E_NCCL, a->err.empty() ? "comm_init: NCCL not available" : a->err.c_str());
  RSC_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) {
    ctx->err = std::string("ncclCommInitRank: ") + (a->GetErrorString ? a->GetErrorString(r) : "error");
    return RSC_E_NCCL;
  }
  ctx->comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->allreduce = nccl_allreduce;
  ctx->allreduce_user = ctx;
  return RSC_OK;
}

int32_t rsc_ctx_allreduce(rsc_ctx* ctx, int32_t* d_buf, int64_t count, void* stream) {
  if (!ctx) return RSC_E_ARG;
  if (!d_buf || count < 0) return fail(ctx, RSC_E_ARG, "allreduce: null buffer");
  if (!ctx->allreduce) return fail(ctx, RSC_E_STATE, "allreduce: the context has no communicator (rsc_ctx_comm_init / rsc_ctx_set_allreduce)");
  if (count == 0) return RSC_OK;
  if (ctx->allreduce(ctx->allreduce_user, d_buf, count, stream ? stream : (void*)ctx->stream))
    return fail(ctx, RSC_E_NCCL, ctx->comm ? ctx->err.c_str() : "allreduce: the callback failed");
  return RSC_OK;
}

int32_t rsc_ctx_comm_stats(rsc_ctx* ctx, int64_t* calls, int64_t* bytes) {
  if (!ctx) return RSC_E_ARG;
  if (calls) *calls = ctx->allreduce_calls;
  if (bytes) *bytes = ctx->allreduce_bytes;
  return RSC_OK;
}

int32_t rsc_ctx_comm_destroy(rsc_ctx* ctx) {
  if (!ctx) return RSC_E_ARG;
  if (!ctx->comm) return RSC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  NcclApi* a = nccl_api();
  if (a->CommDestroy) a->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  if (ctx->allreduce == nccl_allreduce) {
    ctx->allreduce = nullptr;
    ctx->allreduce_user = nullptr;
  }
  ctx->rank = 0, ctx->nranks = 1;
  return RSC_OK;
}

}  // extern "C"
