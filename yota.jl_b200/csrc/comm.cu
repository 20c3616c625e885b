// Multi-GPU: batch-sharded data parallelism.  The tape is replicated, each GPU runs its
// column shard of the batch, and parameter gradients are summed with ncclAllReduce over
// NVLink 5 / NVSwitch on a dedicated stream while the pullback sweep continues (program.cu).
// The reference has no multi-device code at all.
// NCCL is loaded with dlopen so the library also loads where NCCL is absent; any
// communication entry then fails loudly.
#include <dlfcn.h>
#include <limits.h>
#include <stdlib.h>

#include "common.h"

namespace yota {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };
// ncclConfig_t as of NCCL 2.27 (a newer library accepts the older, shorter struct by its
// size / version fields)
struct ncclConfig27 {
  size_t size;
  unsigned int magic, version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs;
  const char* netName;
  int splitShare, trafficClass;
  const char* commName;
  int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
};
constexpr int kUndef = INT_MIN;

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig27*) = nullptr;  // optional
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.h) return YOTA_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) YOTA_FAIL(YOTA_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                              \
  if (!g_nccl.field) YOTA_FAIL(YOTA_E_NCCL, "libnccl is missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  *(void**)(&g_nccl.CommInitRankConfig) = dlsym(g_nccl.h, "ncclCommInitRankConfig");
  return YOTA_OK;
}

#define YOTA_NCCL(expr)                                                                            \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != 0) YOTA_FAIL(YOTA_E_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
  } while (0)

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// SMs the persistent GEMM grids leave free while reductions are in flight: as many as the
// communicator may use CTAs (YOTA_COMM_SM_RESERVE overrides), rounded up to whole CTA pairs.
int comm_reserved_sms(yota_ctx* ctx) {
  if (!ctx->nccl_comm) return 0;
  const int ctas = symm_ctas(ctx) > 0 ? symm_ctas(ctx) : ctx->comm_ctas;
  int r = env_int("YOTA_COMM_SM_RESERVE", ctas);
  if (r < 0) r = 0;
  return (r + 1) & ~1;
}

int comm_allreduce_bucket(yota_ctx* ctx, float* const* bufs, const long long* numel, int n, cudaStream_t s) {
  if (!ctx->nccl_comm) YOTA_FAIL(YOTA_E_NCCL, "yota_comm_init has not been called");
  std::vector<float*> sb, nb;
  std::vector<long long> sn, nn;
  for (int i = 0; i < n; i++) {
    if (numel[i] == 0) continue;
    if (symm_contains(ctx, bufs[i], (size_t)numel[i] * 4)) {
      sb.push_back(bufs[i]);
      sn.push_back(numel[i]);
    } else {
      nb.push_back(bufs[i]);
      nn.push_back(numel[i]);
    }
  }
  if (!sb.empty()) YOTA_TRY(symm_allreduce(ctx, sb.data(), sn.data(), (int)sb.size(), s));
  if (!nb.empty()) {
    YOTA_NCCL(g_nccl.GroupStart());
    for (size_t i = 0; i < nb.size(); i++)
      YOTA_NCCL(g_nccl.AllReduce(nb[i], nb[i], (size_t)nn[i], ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, s));
    YOTA_NCCL(g_nccl.GroupEnd());
  }
  return YOTA_OK;
}

}  // namespace yota

using namespace yota;

extern "C" {

int yota_comm_unique_id(void* id128) {
  if (!id128) YOTA_FAIL(YOTA_E_ARG, "null id buffer");
  YOTA_TRY(nccl_load());
  ncclUniqueId id;
  YOTA_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return YOTA_OK;
}

int yota_comm_init(yota_ctx* ctx, int rank, int nranks, const void* id128) {
  if (!ctx || !id128 || rank < 0 || rank >= nranks) YOTA_FAIL(YOTA_E_ARG, "yota_comm_init: bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(nccl_load());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  // The reductions run beside the tensor-core sweep, so the communicator is capped to a few
  // CTAs (NVSwitch does the arithmetic; 16 CTAs drive the links) and the GEMM grids leave
  // exactly those SMs free (comm_reserved_sms).  YOTA_NCCL_MAX_CTAS=0 keeps NCCL's default.
  const int max_ctas = env_int("YOTA_NCCL_MAX_CTAS", 16);
  int version = 0;
  g_nccl.GetVersion(&version);
  if (max_ctas > 0 && g_nccl.CommInitRankConfig && version >= 22700) {
    ncclConfig27 cfg = {sizeof(ncclConfig27), 0xcafebeefu, 22703u, kUndef, kUndef, kUndef, kUndef, nullptr,
                        kUndef, kUndef, nullptr, kUndef, kUndef, kUndef, kUndef};
    cfg.maxCTAs = max_ctas;
    cfg.minCTAs = env_int("YOTA_NCCL_MIN_CTAS", kUndef);
    cfg.CTAPolicy = env_int("YOTA_NCCL_CTA_POLICY", kUndef);
    cfg.nvlsCTAs = env_int("YOTA_NCCL_NVLS_CTAS", kUndef);
    YOTA_NCCL(g_nccl.CommInitRankConfig(&comm, nranks, id, rank, &cfg));
    ctx->comm_ctas = max_ctas;
  } else {
    YOTA_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm_ctas = 0;
  }
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return YOTA_OK;
}

int yota_comm_destroy(yota_ctx* ctx) {
  if (ctx && ctx->nccl_comm) {
    symm_shutdown(ctx);
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return YOTA_OK;
}

int yota_allreduce_f32(yota_ctx* ctx, float* buf, int64_t n) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  // On the communication stream like the reductions of a run (they share the kernel's barrier
  // slots, so two of them must never be in flight at once), ordered after the compute stream's
  // work so far; the compute stream then waits for it.
  const long long nn = n;
  YOTA_CUDA(cudaEventRecord(ctx->fence_event, ctx->stream));
  YOTA_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->fence_event, 0));
  YOTA_TRY(comm_allreduce_bucket(ctx, &buf, &nn, 1, ctx->comm_stream));
  YOTA_TRY(ctx_add_pending(ctx, {{reinterpret_cast<const char*>(buf), reinterpret_cast<const char*>(buf + n)}}));
  return ctx_join_comm(ctx, ctx->stream);
}

}  // extern "C"
