// All-reduce of parameter gradients over NVSwitch, written against peer / multicast memory
// instead of calling ncclAllReduce: one small kernel per bucket that
//   1. meets the same-index CTA of every other rank on a multicast barrier,
//   2. pulls ITS 1/n slice of the bucket through the switch with multimem.ld_reduce (the
//      NVSwitch adds the n ranks' copies in flight and returns one reduced 16-byte vector),
//   3. pushes the reduced slice back to every rank with multimem.st (one store, n copies),
//   4. meets the others again so that nobody leaves while stores to it are in flight.
// Per GPU and direction that is ~1.125 x the bucket over NVLink -- the floor for an all-reduce
// through an NVLS switch -- with a fixed, small CTA count that the persistent GEMM grids leave
// free (yota_ctx::sm_reserved), so the reductions run beside the pullback sweep.  Without a
// multicast team on the communicator the same kernel reads the n copies through peer pointers
// in rank order and stores to every peer (on this pool 2, 4 and 8 ranks all get multicast).
//
// The buffers must live in symmetric memory (yota_comm_symm_alloc: ncclMemAlloc +
// ncclCommWindowRegister(NCCL_WIN_COLL_SYMMETRIC)); NCCL 2.28's device API supplies the
// multicast / peer address translation and the barrier storage (ncclDevComm).  Buffers
// outside a symmetric window fall back to ncclAllReduce (comm.cu).
//
// Result: every rank holds bit-identical sums (each element is reduced exactly once, by its
// owner rank, and broadcast).  The reference has no multi-device code.
#include <dlfcn.h>
#include <stdlib.h>

#include "common.h"

#if __has_include(<nccl_device.h>)
#include <nccl.h>
#include <nccl_device.h>
#define YOTA_HAVE_NCCL_DEVICE 1
#else
#define YOTA_HAVE_NCCL_DEVICE 0
#endif

namespace yota {

#if YOTA_HAVE_NCCL_DEVICE

struct SymmApi {
  ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
  ncclResult_t (*MemFree)(void*) = nullptr;
  ncclResult_t (*WindowRegister)(ncclComm_t, void*, size_t, ncclWindow_t*, int) = nullptr;
  ncclResult_t (*WindowDeregister)(ncclComm_t, ncclWindow_t) = nullptr;
  ncclResult_t (*DevCommCreate)(ncclComm_t, ncclDevCommRequirements_t const*, ncclDevComm_t*) = nullptr;
  ncclResult_t (*DevCommDestroy)(ncclComm_t, ncclDevComm_t const*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  const char* (*GetLastError)(ncclComm_t) = nullptr;
  bool loaded = false, ok = false;
};
static SymmApi g_symm;

static bool symm_load() {
  if (g_symm.loaded) return g_symm.ok;
  g_symm.loaded = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  *(void**)(&g_symm.MemAlloc) = dlsym(h, "ncclMemAlloc");
  *(void**)(&g_symm.MemFree) = dlsym(h, "ncclMemFree");
  *(void**)(&g_symm.WindowRegister) = dlsym(h, "ncclCommWindowRegister");
  *(void**)(&g_symm.WindowDeregister) = dlsym(h, "ncclCommWindowDeregister");
  *(void**)(&g_symm.DevCommCreate) = dlsym(h, "ncclDevCommCreate");
  *(void**)(&g_symm.DevCommDestroy) = dlsym(h, "ncclDevCommDestroy");
  *(void**)(&g_symm.GetErrorString) = dlsym(h, "ncclGetErrorString");
  *(void**)(&g_symm.GetLastError) = dlsym(h, "ncclGetLastError");
  g_symm.ok = g_symm.MemAlloc && g_symm.MemFree && g_symm.WindowRegister && g_symm.WindowDeregister &&
              g_symm.DevCommCreate && g_symm.DevCommDestroy && g_symm.GetErrorString;
  return g_symm.ok;
}

#define YOTA_NCCL2(expr)                                                                              \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess)                                                                            \
      YOTA_FAIL(YOTA_E_NCCL, "NCCL error at %s:%d: %s (%s)", __FILE__, __LINE__, g_symm.GetErrorString(_r), \
                g_symm.GetLastError ? g_symm.GetLastError(nullptr) : "");                             \
  } while (0)

constexpr int AR_MAX_SEG = 8;
// 256 threads x 8 vectors in flight: enough to keep NVLink busy (the reduction is link-bound at
// ~180 us per 64 MiB on 8 GPUs), few enough that the requests do not crowd the GEMM's TMA loads
// out of the memory system -- with 1024 threads the overlapped GEMM stretched to the length of
// the reduction (measured, tools/gpu_ar_diag.sh; 8 GPUs: 415 -> 476 evals/s).
constexpr int AR_THREADS = 256;
constexpr int AR_UNROLL = 8;

struct ArSegs {
  int n;
  ncclWindow_t win[AR_MAX_SEG];
  size_t off[AR_MAX_SEG];  // byte offset of the buffer inside its window
  long long numel[AR_MAX_SEG];
};

struct SymmState {
  ncclDevComm dev{};
  bool have_dev = false, multimem = false;
  int ctas = 0;
  struct Win {
    char* base;
    size_t size;
    ncclWindow_t win;
  };
  std::vector<Win> wins;
};

__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float mc_ld_reduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Launched as clusters of 2 CTAs: a cluster of two lands on the two SMs of ONE TPC, so k CTAs
// take k/2 whole TPCs and leave the others whole for the GEMM's CTA pairs (cta_group::2 needs
// both SMs of a TPC; CTAs sprinkled one per TPC would block that many pairs for the whole
// reduction -- measured: the overlapped dh GEMM went from 62 us to 170 us).
template <bool MC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1) symm_allreduce_kernel(const ncclDevComm dc, const ArSegs segs, const int phases) {
  ncclLsaBarrierSession<ncclCoopCta> bar(ncclCoopCta(), dc, ncclTeamTagLsa(), blockIdx.x, MC);
  // every rank's producer kernel (the dW GEMM) has finished once its CTA arrives here
  bar.sync(ncclCoopCta(), cuda::memory_order_acq_rel);
  const int nr = dc.lsaSize, r = dc.lsaRank;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (int sidx = 0; sidx < segs.n; sidx++) {
    const long long numel = segs.numel[sidx];
    const long long nv = numel >> 2;
    const long long lo = nv * r / nr, hi = nv * (r + 1) / nr;  // this rank's slice, in float4 units
    if (MC) {
      float4* mm = reinterpret_cast<float4*>(ncclGetLsaMultimemPointer(segs.win[sidx], segs.off[sidx], dc));
      for (long long i0 = lo + tid; i0 < hi; i0 += nthreads * AR_UNROLL) {
        float4 v[AR_UNROLL] = {};
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
          const long long i = i0 + (long long)u * nthreads;
          if (i < hi && (phases & 1)) v[u] = mc_ld_reduce(mm + i);
        }
#pragma unroll
        for (int u = 0; u < AR_UNROLL; u++) {
          const long long i = i0 + (long long)u * nthreads;
          if (i < hi && (phases & 2)) mc_st(mm + i, v[u]);
        }
      }
      if (r == 0 && tid < (numel & 3)) {  // ragged tail (numel not a multiple of 4): scalar form
        float* m1 = reinterpret_cast<float*>(mm) + (nv << 2) + tid;
        mc_st1(m1, mc_ld_reduce1(m1));
      }
    } else {
      for (long long i0 = lo + tid; i0 < hi; i0 += nthreads * 2) {
        float4 acc[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const long long i = i0 + (long long)u * nthreads;
          if (i < hi) {
            acc[u] = ld_sys(reinterpret_cast<const float4*>(ncclGetLsaPointer(segs.win[sidx], segs.off[sidx], 0)) + i);
            for (int p = 1; p < nr; p++) {  // rank order: the same sum on every run
              const float4 w =
                  ld_sys(reinterpret_cast<const float4*>(ncclGetLsaPointer(segs.win[sidx], segs.off[sidx], p)) + i);
              acc[u].x += w.x;
              acc[u].y += w.y;
              acc[u].z += w.z;
              acc[u].w += w.w;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const long long i = i0 + (long long)u * nthreads;
          if (i < hi)
            for (int p = 0; p < nr; p++)
              st_sys(reinterpret_cast<float4*>(ncclGetLsaPointer(segs.win[sidx], segs.off[sidx], p)) + i, acc[u]);
        }
      }
      if (r == 0 && tid < (numel & 3)) {
        const long long i = (nv << 2) + tid;
        float acc = 0.f;
        for (int p = 0; p < nr; p++) {
          const float* src = reinterpret_cast<const float*>(ncclGetLsaPointer(segs.win[sidx], segs.off[sidx], p)) + i;
          float w;
          asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(w) : "l"(src) : "memory");
          acc = p == 0 ? w : acc + w;
        }
        for (int p = 0; p < nr; p++) {
          float* dst = reinterpret_cast<float*>(ncclGetLsaPointer(segs.win[sidx], segs.off[sidx], p)) + i;
          asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(acc) : "memory");
        }
      }
    }
  }
  // nobody leaves (and lets its next kernel touch the buffers) before every rank's stores landed
  bar.sync(ncclCoopCta(), cuda::memory_order_acq_rel);
}

static int env_int2(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static SymmState* symm_state(yota_ctx* ctx) { return static_cast<SymmState*>(ctx->symm); }

static int symm_ensure_devcomm(yota_ctx* ctx) {
  if (!ctx->symm) ctx->symm = new SymmState();
  SymmState* S = symm_state(ctx);
  if (S->have_dev) return YOTA_OK;
  S->ctas = (env_int2("YOTA_AR_CTAS", 16) + 1) & ~1;
  if (S->ctas < 2) S->ctas = 2;
  ncclDevCommRequirements reqs;
  memset(&reqs, 0, sizeof(reqs));
  reqs.lsaBarrierCount = S->ctas;
  reqs.lsaMultimem = !getenv("YOTA_AR_NO_MULTIMEM");
  ncclResult_t r = g_symm.DevCommCreate((ncclComm_t)ctx->nccl_comm, &reqs, &S->dev);
  if (r != ncclSuccess && reqs.lsaMultimem) {  // no multicast team on this communicator: peer pointers
    reqs.lsaMultimem = false;
    r = g_symm.DevCommCreate((ncclComm_t)ctx->nccl_comm, &reqs, &S->dev);
  }
  if (r != ncclSuccess)
    YOTA_FAIL(YOTA_E_NCCL, "ncclDevCommCreate failed: %s", g_symm.GetErrorString(r));
  S->multimem = reqs.lsaMultimem && S->dev.lsaMultimem.mcBasePtr != nullptr;
  if (S->dev.lsaSize != ctx->nranks)
    YOTA_FAIL(YOTA_E_NCCL, "symmetric all-reduce needs every rank load/store accessible (lsa team %d of %d ranks)",
              S->dev.lsaSize, ctx->nranks);
  S->have_dev = true;
  return YOTA_OK;
}

bool symm_available(yota_ctx* ctx) { return ctx->nccl_comm && symm_load() && !getenv("YOTA_AR_NCCL"); }

int symm_alloc(yota_ctx* ctx, size_t bytes, void** out) {
  if (!symm_available(ctx)) YOTA_FAIL(YOTA_E_NCCL, "symmetric memory needs yota_comm_init and NCCL >= 2.28");
  YOTA_TRY(symm_ensure_devcomm(ctx));
  SymmState* S = symm_state(ctx);
  bytes = (bytes + 4095) & ~(size_t)4095;
  void* p = nullptr;
  YOTA_NCCL2(g_symm.MemAlloc(&p, bytes));
  ncclWindow_t w = nullptr;
  ncclResult_t r = g_symm.WindowRegister((ncclComm_t)ctx->nccl_comm, p, bytes, &w, NCCL_WIN_COLL_SYMMETRIC);
  if (r != ncclSuccess || !w) {
    g_symm.MemFree(p);
    YOTA_FAIL(YOTA_E_NCCL, "ncclCommWindowRegister failed: %s", g_symm.GetErrorString(r));
  }
  S->wins.push_back({static_cast<char*>(p), bytes, w});
  *out = p;
  return YOTA_OK;
}

int symm_free(yota_ctx* ctx, void* ptr) {
  SymmState* S = ctx ? symm_state(ctx) : nullptr;
  if (!S) return YOTA_OK;
  for (size_t i = 0; i < S->wins.size(); i++)
    if (S->wins[i].base == ptr) {
      if (ctx->nccl_comm) g_symm.WindowDeregister((ncclComm_t)ctx->nccl_comm, S->wins[i].win);
      g_symm.MemFree(ptr);
      S->wins.erase(S->wins.begin() + i);
      return YOTA_OK;
    }
  YOTA_FAIL(YOTA_E_ARG, "yota_comm_symm_free: not a symmetric allocation");
}

void symm_shutdown(yota_ctx* ctx) {
  SymmState* S = ctx ? symm_state(ctx) : nullptr;
  if (!S) return;
  while (!S->wins.empty()) symm_free(ctx, S->wins.back().base);
  if (S->have_dev && ctx->nccl_comm) g_symm.DevCommDestroy((ncclComm_t)ctx->nccl_comm, &S->dev);
  delete S;
  ctx->symm = nullptr;
}

bool symm_contains(yota_ctx* ctx, const void* ptr, size_t bytes) {
  SymmState* S = symm_state(ctx);
  if (!S || !S->have_dev) return false;
  const char* p = static_cast<const char*>(ptr);
  for (auto& w : S->wins)
    if (p >= w.base && p + bytes <= w.base + w.size) return true;
  return false;
}

int symm_ctas(yota_ctx* ctx) {
  SymmState* S = symm_state(ctx);
  return S && S->have_dev ? S->ctas : 0;
}
bool symm_multimem(yota_ctx* ctx) {
  SymmState* S = symm_state(ctx);
  return S && S->have_dev && S->multimem;
}

// one kernel for up to AR_MAX_SEG buffers (a layer's weight and bias gradients travel together)
int symm_allreduce(yota_ctx* ctx, float* const* bufs, const long long* numel, int n, cudaStream_t s) {
  SymmState* S = symm_state(ctx);
  if (!S || !S->have_dev) YOTA_FAIL(YOTA_E_NCCL, "symmetric all-reduce without symmetric memory");
  for (int b0 = 0; b0 < n; b0 += AR_MAX_SEG) {
    ArSegs segs;
    memset(&segs, 0, sizeof(segs));
    long long total = 0;
    for (int b = b0; b < n && b < b0 + AR_MAX_SEG; b++) {
      const char* p = reinterpret_cast<const char*>(bufs[b]);
      const SymmState::Win* win = nullptr;
      for (auto& w : S->wins)
        if (p >= w.base && p + numel[b] * 4 <= w.base + w.size) win = &w;
      if (!win) YOTA_FAIL(YOTA_E_ARG, "symm_allreduce: buffer is not in symmetric memory");
      if ((p - win->base) & 15) YOTA_FAIL(YOTA_E_ARG, "symm_allreduce: buffer is not 16-byte aligned");
      segs.win[segs.n] = win->win;
      segs.off[segs.n] = (size_t)(p - win->base);
      segs.numel[segs.n] = numel[b];
      segs.n++;
      total += numel[b];
    }
    // small buckets: one CTA per 64 KiB of slice is plenty; large ones use every reserved CTA
    int ctas = S->ctas;
    const long long per_rank_vec = total / 4 / ctx->nranks + 1;
    const long long want = (per_rank_vec + AR_THREADS * AR_UNROLL - 1) / (AR_THREADS * AR_UNROLL);
    // (sized before the diagnostic thread-count override below: only large buckets are affected)
    if (want < ctas) ctas = (int)(want < 1 ? 1 : want);
    ctas = (ctas + 1) & ~1;  // whole clusters
    const int phases = env_int2("YOTA_AR_PHASES", 3);    // diagnostics: 1 = reduce only, 2 = broadcast only, 0 = barriers only
    const int threads = env_int2("YOTA_AR_THREADS", AR_THREADS);
    if (S->multimem)
      symm_allreduce_kernel<true><<<ctas, threads, 0, s>>>(S->dev, segs, phases);
    else
      symm_allreduce_kernel<false><<<ctas, threads, 0, s>>>(S->dev, segs, phases);
    YOTA_CUDA(cudaGetLastError());
  }
  return YOTA_OK;
}

#else  // !YOTA_HAVE_NCCL_DEVICE

bool symm_available(yota_ctx*) { return false; }
int symm_alloc(yota_ctx*, size_t, void**) { YOTA_FAIL(YOTA_E_NCCL, "built without the NCCL device API (nccl_device.h)"); }
int symm_free(yota_ctx*, void*) { return YOTA_OK; }
void symm_shutdown(yota_ctx*) {}
bool symm_contains(yota_ctx*, const void*, size_t) { return false; }
int symm_ctas(yota_ctx*) { return 0; }
bool symm_multimem(yota_ctx*) { return false; }
int symm_allreduce(yota_ctx*, float* const*, const long long*, int, cudaStream_t) {
  YOTA_FAIL(YOTA_E_NCCL, "built without the NCCL device API (nccl_device.h)");
}

#endif

}  // namespace yota

using namespace yota;

extern "C" {

int yota_comm_symm_alloc(yota_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) YOTA_FAIL(YOTA_E_ARG, "yota_comm_symm_alloc: null argument");
  YOTA_TRY(ctx_ensure_device(ctx));
  return symm_alloc(ctx, bytes, dptr);
}

int yota_comm_symm_free(yota_ctx* ctx, void* dptr) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  return symm_free(ctx, dptr);
}

int yota_comm_symm_available(yota_ctx* ctx) {
  if (!ctx || !symm_available(ctx)) return 0;
  return symm_multimem(ctx) ? 2 : 1;
}

}  // extern "C"
