// C ABI: context, caller-owned device buffers, events, stand-alone operators.
#include <stdarg.h>

#include <algorithm>
#include <string.h>

#include "common.h"

namespace yota {

static thread_local char g_err[1024] = "";
thread_local int g_last_status = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ctx_ensure_device(yota_ctx* ctx) {
  if (ctx->device_ready) {
    YOTA_CUDA(cudaSetDevice(ctx->device));
    return YOTA_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    YOTA_FAIL(YOTA_E_CUDA, "no CUDA device available (%s): libyota_cuda has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (ctx->device >= n) YOTA_FAIL(YOTA_E_ARG, "device %d requested but only %d present", ctx->device, n);
  YOTA_CUDA(cudaSetDevice(ctx->device));
  cudaDeviceProp pr;
  YOTA_CUDA(cudaGetDeviceProperties(&pr, ctx->device));
  ctx->sm_count = pr.multiProcessorCount;
  ctx->cc_major = pr.major;
  ctx->cc_minor = pr.minor;
  ctx->total_mem = pr.totalGlobalMem;
  if (pr.major != 10)
    YOTA_FAIL(YOTA_E_CUDA, "device %d is sm_%d%d; libyota_cuda is built for sm_100a (B200) only", ctx->device, pr.major,
              pr.minor);
  YOTA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  YOTA_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  YOTA_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  YOTA_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  for (auto& ls : ctx->leaf_streams) YOTA_CUDA(cudaStreamCreateWithFlags(&ls, cudaStreamNonBlocking));
  YOTA_CUDA(cudaEventCreateWithFlags(&ctx->fence_event, cudaEventDisableTiming));
  YOTA_CUDA(cudaEventCreateWithFlags(&ctx->free_event, cudaEventDisableTiming));
  YOTA_CUDA(cudaHostAlloc((void**)&ctx->err_host, 4 * sizeof(long long), cudaHostAllocMapped));
  memset(ctx->err_host, 0, 4 * sizeof(long long));
  YOTA_CUDA(cudaHostGetDevicePointer((void**)&ctx->err_dev, ctx->err_host, 0));
  ctx->device_ready = true;
  return YOTA_OK;
}

static bool overlaps(const yota_ctx::PendingAr& e, const char* lo, const char* hi) {
  for (auto& r : e.ranges)
    if (r.first < hi && lo < r.second) return true;
  return false;
}

int ctx_wait_pending(yota_ctx* ctx, const void* ptr, size_t bytes, cudaStream_t s) {
  if (ctx->pending_ar.empty() || !ptr) return YOTA_OK;
  const char* lo = static_cast<const char*>(ptr);
  for (auto& e : ctx->pending_ar)
    if (overlaps(e, lo, lo + (bytes ? bytes : 1))) YOTA_CUDA(cudaStreamWaitEvent(s, e.ev, 0));
  return YOTA_OK;
}

int ctx_join_comm(yota_ctx* ctx, cudaStream_t s) {
  // the communication stream is in order: the newest event implies all older ones
  if (!ctx->pending_ar.empty()) YOTA_CUDA(cudaStreamWaitEvent(s, ctx->pending_ar.back().ev, 0));
  for (auto& e : ctx->pending_ar) ctx->event_pool.push_back(e.ev);
  ctx->pending_ar.clear();
  return YOTA_OK;
}

int ctx_add_pending(yota_ctx* ctx, const std::vector<std::pair<const char*, const char*>>& ranges) {
  yota_ctx::PendingAr n;
  n.ranges = ranges;
  // an older entry over the same bytes is superseded (same stream, later event); bytes it covered
  // beyond the new ranges ride along with the new, later event (waits longer, never too short)
  for (size_t i = 0; i < ctx->pending_ar.size();) {
    bool hit = false;
    for (auto& r : ranges) hit = hit || overlaps(ctx->pending_ar[i], r.first, r.second);
    if (!hit) {
      i++;
      continue;
    }
    for (auto& r : ctx->pending_ar[i].ranges)
      if (std::find(n.ranges.begin(), n.ranges.end(), r) == n.ranges.end()) n.ranges.push_back(r);
    ctx->event_pool.push_back(ctx->pending_ar[i].ev);
    ctx->pending_ar.erase(ctx->pending_ar.begin() + i);
  }
  if (ctx->event_pool.empty()) {
    cudaEvent_t e;
    YOTA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->event_pool.push_back(e);
  }
  n.ev = ctx->event_pool.back();
  ctx->event_pool.pop_back();
  YOTA_CUDA(cudaEventRecord(n.ev, ctx->comm_stream));
  ctx->pending_ar.push_back(n);
  return YOTA_OK;
}

int ctx_check_index_errors(yota_ctx* ctx) {
  if (!ctx->err_host) return YOTA_OK;
  if (*(volatile long long*)(ctx->err_host + 3)) {
    memset(ctx->err_host, 0, 4 * sizeof(long long));
    YOTA_FAIL(YOTA_E_CUDA, "internal: a persistent recurrence kernel gave up waiting for a tile of its previous step "
                           "(the results of that run are invalid; YOTA_NO_CHAIN=1 launches the steps one by one)");
  }
  if (!*(volatile long long*)ctx->err_host) return YOTA_OK;
  const long long j = ctx->err_host[1], n = ctx->err_host[2];
  memset(ctx->err_host, 0, 4 * sizeof(long long));
  YOTA_FAIL(YOTA_E_SHAPE,
            "BoundsError: attempt to access a dimension of extent %lld at index [%lld] (run-time index vector of a "
            "getindex / scatter-add; the results of that run are invalid)",
            n, j);
}

int ctx_scratch(yota_ctx* ctx, size_t bytes, void** out) {
  if (ctx->scratch_bytes < bytes) {
    if (ctx->scratch) {
      YOTA_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->scratch);
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->scratch, bytes);
    if (e != cudaSuccess) YOTA_FAIL(YOTA_E_NOMEM, "cannot allocate %zu bytes of scratch: %s", bytes, cudaGetErrorString(e));
    ctx->scratch_bytes = bytes;
  }
  *out = ctx->scratch;
  return YOTA_OK;
}

}  // namespace yota

using namespace yota;

extern "C" {

int yota_abi_version(void) { return YOTA_ABI_VERSION; }
const char* yota_last_error(void) { return g_err; }

int yota_init(int device, yota_ctx** out) {
  if (!out || device < 0) YOTA_FAIL(YOTA_E_ARG, "yota_init: bad arguments");
  yota_ctx* c = new yota_ctx();
  c->device = device;
  // The device itself is attached on first use so that planning (yota_program_build /
  // describe) also works on a host without a GPU; every compute entry fails loudly there.
  *out = c;
  return YOTA_OK;
}

int yota_shutdown(yota_ctx* ctx) {
  if (!ctx) return YOTA_OK;
  if (ctx->device_ready) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    yota_comm_destroy(ctx);
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    for (auto& kv : ctx->live_blocks) cudaFree(kv.first);
    if (ctx->scratch) cudaFree(ctx->scratch);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->comm_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->side_stream);
    for (auto ls : ctx->leaf_streams) cudaStreamDestroy(ls);
    cudaEventDestroy(ctx->fence_event);
    if (ctx->free_event) cudaEventDestroy(ctx->free_event);
    for (auto& e : ctx->pending_ar) cudaEventDestroy(e.ev);
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->err_host) cudaFreeHost(ctx->err_host);
  }
  delete ctx;
  return YOTA_OK;
}

int yota_device_info(yota_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (total_mem) *total_mem = ctx->total_mem;
  return YOTA_OK;
}

// Pooled allocation: grad() hands out fresh result arrays every call (Julia semantics), so
// freed blocks are kept in exact-size buckets and reused without touching cudaMalloc.
int yota_alloc(yota_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) YOTA_FAIL(YOTA_E_ARG, "yota_alloc: bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  bytes = (bytes + 511) & ~(size_t)511;
  if (bytes == 0) bytes = 512;
  auto it = ctx->free_blocks.find(bytes);
  if (it != ctx->free_blocks.end()) {
    *dptr = it->second;
    ctx->free_blocks.erase(it);
  } else {
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) {
      // drop the cache and retry once
      for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
      ctx->free_blocks.clear();
      e = cudaMalloc(dptr, bytes);
      if (e != cudaSuccess) YOTA_FAIL(YOTA_E_NOMEM, "yota_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
  }
  ctx->live_blocks[*dptr] = bytes;
  return YOTA_OK;
}

int yota_free(yota_ctx* ctx, void* dptr) {
  if (!ctx || !dptr) return YOTA_OK;
  auto it = ctx->live_blocks.find(dptr);
  if (it == ctx->live_blocks.end()) YOTA_FAIL(YOTA_E_ARG, "yota_free: pointer was not allocated by yota_alloc");
  // stream-ordered reuse: consumers on ctx->stream touch a recycled block only after the work
  // already queued on it; the COPY stream is ordered behind that work the next time it is used
  // (yota_h2d_async), the communication stream always waits for the compute stream before a
  // reduction (program.cu)
  ctx->free_pending = true;
  YOTA_TRY(ctx_wait_pending(ctx, dptr, it->second, ctx->stream));  // a reduction may still be writing it
  ctx->free_blocks.insert({it->second, dptr});
  ctx->live_blocks.erase(it);
  return YOTA_OK;
}

int yota_h2d(yota_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_wait_pending(ctx, dst, bytes, ctx->stream));
  YOTA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  YOTA_CUDA(cudaStreamSynchronize(ctx->stream));
  return YOTA_OK;
}

int yota_d2h(yota_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_wait_pending(ctx, src, bytes, ctx->stream));  // the value of a reduced buffer is the reduced one
  YOTA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  YOTA_CUDA(cudaStreamSynchronize(ctx->stream));
  return ctx_check_index_errors(ctx);
}

int yota_d2h_async(yota_ctx* ctx, void* dst_pinned, const void* src, size_t bytes) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_wait_pending(ctx, src, bytes, ctx->stream));
  YOTA_CUDA(cudaMemcpyAsync(dst_pinned, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return YOTA_OK;
}

int yota_host_alloc(yota_ctx* ctx, size_t bytes, void** hptr) {
  if (!ctx || !hptr) YOTA_FAIL(YOTA_E_ARG, "yota_host_alloc: bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
  return YOTA_OK;
}

int yota_host_free(yota_ctx* ctx, void* hptr) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  if (hptr) YOTA_CUDA(cudaFreeHost(hptr));
  return YOTA_OK;
}

int yota_h2d_async(yota_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  if (ctx->free_pending) {  // dst may be a block recycled from kernels still queued on the compute stream
    YOTA_CUDA(cudaEventRecord(ctx->free_event, ctx->stream));
    YOTA_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->free_event, 0));
    ctx->free_pending = false;
  }
  YOTA_TRY(ctx_wait_pending(ctx, dst, bytes, ctx->copy_stream));
  YOTA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  return YOTA_OK;
}

int yota_stream_fence(yota_ctx* ctx, int from, int to) {
  if (!ctx || from == to || from < 0 || from > 1 || to < 0 || to > 1) YOTA_FAIL(YOTA_E_ARG, "yota_stream_fence: bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  cudaStream_t s[2] = {ctx->stream, ctx->copy_stream};
  YOTA_CUDA(cudaEventRecord(ctx->fence_event, s[from]));
  YOTA_CUDA(cudaStreamWaitEvent(s[to], ctx->fence_event, 0));
  return YOTA_OK;
}

int yota_memset(yota_ctx* ctx, void* dptr, int byte, size_t bytes) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_wait_pending(ctx, dptr, bytes, ctx->stream));
  YOTA_CUDA(cudaMemsetAsync(dptr, byte, bytes, ctx->stream));
  return YOTA_OK;
}

int yota_sync(yota_ctx* ctx) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_CUDA(cudaStreamSynchronize(ctx->stream));
  YOTA_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  YOTA_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  for (auto& e : ctx->pending_ar) ctx->event_pool.push_back(e.ev);
  ctx->pending_ar.clear();
  return ctx_check_index_errors(ctx);
}

int yota_event_create(yota_ctx* ctx, void** ev) {
  if (!ctx || !ev) YOTA_FAIL(YOTA_E_ARG, "bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  cudaEvent_t e;
  YOTA_CUDA(cudaEventCreate(&e));
  *ev = e;
  return YOTA_OK;
}
int yota_event_record(yota_ctx* ctx, void* ev) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  // a timing event covers the reductions of the runs issued before it
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  YOTA_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ev), ctx->stream));
  return YOTA_OK;
}
int yota_event_elapsed_ms(yota_ctx* ctx, void* a, void* b, float* ms) {
  (void)ctx;
  YOTA_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(b)));
  YOTA_CUDA(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
  return YOTA_OK;
}
int yota_event_destroy(yota_ctx* ctx, void* ev) {
  (void)ctx;
  if (ev) YOTA_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return YOTA_OK;
}

int yota_gemm(yota_ctx* ctx, int mode, int transA, int transB, int64_t m, int64_t n, int64_t k, const float* A,
              int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int accumulate) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  GemmArgs g{mode, transA, transB, m, n, k, A, lda, B, ldb, C, ldc, accumulate};
  if (mode != 0) {
    if (!gemm_tc_eligible(g))
      YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "yota_gemm: shape %lldx%lldx%lld is not eligible for the tensor-core kernel",
                (long long)m, (long long)n, (long long)k);
    if (mode == 2) {  // BF16 operands: cast shadows of A and B into scratch first
      const long long na = lda * (transA ? m : k), nb = ldb * (transB ? k : n);
      const size_t offb = ((size_t)na * 2 + 255) & ~(size_t)255;
      void* w = nullptr;
      YOTA_TRY(ctx_scratch(ctx, offb + (size_t)nb * 2 + 256, &w));
      YOTA_TRY(launch_cast_bf16(w, A, na, ctx->stream));
      YOTA_TRY(launch_cast_bf16(static_cast<char*>(w) + offb, B, nb, ctx->stream));
      g.A = w;
      g.B = static_cast<char*>(w) + offb;
    }
    if (mode == 3) {  // compensated: lo parts of A and B into scratch, hi is read from the arrays themselves
      const long long na = lda * (transA ? m : k), nb = ldb * (transB ? k : n);
      const size_t offb = ((size_t)na * 4 + 255) & ~(size_t)255;
      void* w = nullptr;
      YOTA_TRY(ctx_scratch(ctx, offb + (size_t)nb * 4 + 256, &w));
      YOTA_TRY(launch_split_tf32(static_cast<float*>(w), A, na, ctx->stream));
      YOTA_TRY(launch_split_tf32(reinterpret_cast<float*>(static_cast<char*>(w) + offb), B, nb, ctx->stream));
      g.A_lo = w;
      g.B_lo = static_cast<char*>(w) + offb;
    }
    return launch_gemm_tc(ctx, g, ctx->stream);
  }
  return launch_gemm_ffma(g, ctx->stream);
}

int yota_scatter_add_cols(yota_ctx* ctx, float* dx, int64_t rows, int64_t n_dst, const float* dy, const int64_t* idx,
                          int64_t n_src) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  const size_t ws = scatter_cols_workspace_bytes(n_dst, n_src);
  void* w = nullptr;
  YOTA_TRY(ctx_scratch(ctx, ws, &w));
  return launch_scatter_add_cols(dx, rows, n_dst, dy, idx, n_src, w, ws, ctx->stream, nullptr, ctx->err_dev);
}

int yota_scatter_add_cols_sharded(yota_ctx* ctx, float* dx_shard, int64_t rows, int64_t n_dst, int64_t dst_lo,
                                  int64_t dst_hi, const float* dy, const int64_t* idx, int64_t n_src) {
  if (!ctx || dst_lo < 0 || dst_hi < dst_lo || dst_hi > n_dst) YOTA_FAIL(YOTA_E_ARG, "yota_scatter_add_cols_sharded: bad destination range");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  const size_t ws = scatter_cols_workspace_bytes(dst_hi - dst_lo, n_src);
  void* w = nullptr;
  YOTA_TRY(ctx_scratch(ctx, ws, &w));
  return launch_scatter_add_cols(dx_shard, rows, dst_hi - dst_lo, dy, idx, n_src, w, ws, ctx->stream, nullptr,
                                 ctx->err_dev, dst_lo, n_dst);
}

int yota_gather_cols(yota_ctx* ctx, float* y, const float* x, int64_t rows, int64_t n_dst, const int64_t* idx,
                     int64_t n_src) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  return launch_gather_cols(y, x, rows, n_dst, idx, n_src, ctx->stream, ctx->err_dev);
}

int yota_update_sgd(yota_ctx* ctx, float* x, const float* gx, int64_t n, float lr) {
  if (!ctx) YOTA_FAIL(YOTA_E_ARG, "null ctx");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  return launch_axpy(x, gx, n, lr, ctx->stream);
}

int yota_update_adam(yota_ctx* ctx, float* x, const float* gx, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, int t) {
  if (!ctx || t < 1) YOTA_FAIL(YOTA_E_ARG, "yota_update_adam: bad arguments");
  YOTA_TRY(ctx_ensure_device(ctx));
  YOTA_TRY(ctx_join_comm(ctx, ctx->stream));
  return launch_adam(x, gx, m, v, n, lr, beta1, beta2, eps, nullptr, t, ctx->stream);
}

}  // extern "C"
