// ctx.cu — context lifetime, NCCL communicator plumbing, library self-description.
#include "lld_ctx.h"

#ifdef LLD_WITH_NCCL
#include <nccl.h>
#endif

struct BaState;
void lld_ba_state_free(BaState*);

extern "C" const char* lld_version(void) { return "lldba 0.1 (sm_100a)"; }

extern "C" int lld_ctx_create(int device, void** out) {
  if (!out) return LLD_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return LLD_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return LLD_ERR_CUDA;
  LldCtx* c = new LldCtx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return LLD_ERR_CUDA; }
  for (int i = 0; i < 4; i++)
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { delete c; return LLD_ERR_CUDA; }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  c->pinned_cap = 1 << 16;
  if (cudaMallocHost(&c->pinned, c->pinned_cap) != cudaSuccess) { delete c; return LLD_ERR_CUDA; }
  *out = c;
  return LLD_OK;
}

extern "C" void lld_ctx_destroy(void* ctx) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
#ifdef LLD_WITH_NCCL
  if (c->comm) ncclCommDestroy(reinterpret_cast<ncclComm_t>(c->comm));
#endif
  for (auto& b : c->pool) b.release();
  if (c->pinned) cudaFreeHost(c->pinned);
  for (int i = 0; i < 4; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  lld_ba_state_free(c->ba);
  delete c;
}

extern "C" const char* lld_ctx_last_error(void* ctx) {
  LldCtx* c = lld_ctx_cast(ctx);
  return c ? c->err : "null context";
}
extern "C" int64_t lld_ctx_launch_count(void* ctx) {
  LldCtx* c = lld_ctx_cast(ctx);
  return c ? c->launches : 0;
}
extern "C" void lld_ctx_last_timing(void* ctx, float* a, float* b, float* d) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c) return;
  if (a) *a = c->ms_h2d;
  if (b) *b = c->ms_compute;
  if (d) *d = c->ms_d2h;
}
extern "C" void* lld_ctx_stream(void* ctx) {
  LldCtx* c = lld_ctx_cast(ctx);
  return c ? (void*)c->stream : nullptr;
}

extern "C" int lld_comm_unique_id(uint8_t id_out[128]) {
#ifdef LLD_WITH_NCCL
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (ncclGetUniqueId(&id) != ncclSuccess) return LLD_ERR_NCCL;
  memcpy(id_out, &id, 128);
  return LLD_OK;
#else
  (void)id_out;
  return LLD_ERR_UNSUPPORTED;
#endif
}

extern "C" int lld_comm_init(void* ctx, int n_ranks, int rank, const uint8_t unique_id[128]) {
  LldCtx* c = lld_ctx_cast(ctx);
  if (!c) return LLD_ERR_ARG;
#ifdef LLD_WITH_NCCL
  LLD_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, unique_id, 128);
  ncclComm_t comm;
  ncclResult_t r = ncclCommInitRank(&comm, n_ranks, id, rank);
  if (r != ncclSuccess) {
    snprintf(c->err, sizeof(c->err), "ncclCommInitRank: %s", ncclGetErrorString(r));
    return LLD_ERR_NCCL;
  }
  c->comm = reinterpret_cast<ncclComm*>(comm);
  c->n_ranks = n_ranks;
  c->rank = rank;
  return LLD_OK;
#else
  (void)n_ranks; (void)rank; (void)unique_id;
  return LLD_ERR_UNSUPPORTED;
#endif
}
