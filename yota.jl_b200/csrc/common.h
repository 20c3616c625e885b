// Internal declarations shared by the translation units of libyota_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/yota_cuda.h"

namespace yota {

void set_error(const char* fmt, ...);
extern thread_local int g_last_status;

#define YOTA_FAIL(code, ...)            \
  do {                                  \
    ::yota::set_error(__VA_ARGS__);     \
    return (code);                      \
  } while (0)

#define YOTA_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::yota::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,     \
                        __LINE__, cudaGetErrorString(_e));                                \
      return YOTA_E_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define YOTA_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != YOTA_OK) return _s; \
  } while (0)

struct NcclApi;  // comm.cu

}  // namespace yota

struct yota_ctx {
  int device = 0;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
  bool device_ready = false;  // lazily initialised so that plan-only use works without a GPU
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // work of a run that depends on its inputs only (index sorts)
  cudaStream_t leaf_streams[3] = {nullptr, nullptr, nullptr};  // small leaf steps of a run (plan_side_leaves)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t fence_event = nullptr;
  // size-bucketed free lists (caller-owned buffers come and go every grad() call)
  std::multimap<size_t, void*> free_blocks;
  std::map<void*, size_t> live_blocks;
  // NCCL
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // SMs left to the communication kernels while all-reduces overlap the pullback sweep: the
  // persistent GEMM grids shrink by this many so that NCCL's CTAs find free SMs at once
  // instead of waiting for (and then delaying) a GEMM wave.  Set by run_all_steps.
  int sm_reserved = 0;
  int comm_ctas = 0;  // CTAs the communicator was configured with (0: NCCL default)
  void* symm = nullptr;  // symmetric-memory state of the NVSwitch all-reduce (symm_allreduce.cu)
  // scratch for standalone ops
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // Run-time index errors (Julia: BoundsError thrown by getindex / NNlib.scatter!,
  // src/helpers.jl:7-11): kernels that read index VECTORS report an out-of-range entry here
  // (pinned, device-mapped: {flag, offending 1-based index, extent}); checked at the next
  // synchronising entry point (yota_sync / yota_d2h), which then fails with YOTA_E_SHAPE.
  long long* err_host = nullptr;
  long long* err_dev = nullptr;
  // caller-owned blocks recycled by yota_free while work on another stream may still touch them
  cudaEvent_t free_event = nullptr;
  bool free_pending = false;
  // Reductions still in flight on the communication stream when yota_program_run returned:
  // the run does NOT join the two streams, so the tail buckets of eval i overlap the forward
  // sweep of eval i+1.  Every later operation that touches one of these byte ranges first
  // makes its stream wait for the bucket's event (ctx_wait_pending); yota_sync,
  // yota_event_record and the stand-alone operators join everything (ctx_join_comm).
  struct PendingAr {
    std::vector<std::pair<const char*, const char*>> ranges;
    cudaEvent_t ev;
  };
  std::vector<PendingAr> pending_ar;
  std::vector<cudaEvent_t> event_pool;
};

namespace yota {

int ctx_ensure_device(yota_ctx* ctx);
// fails with YOTA_E_SHAPE ("BoundsError ...") if a kernel reported an out-of-range run-time index
// since the last check; call after the stream has been synchronised
int ctx_check_index_errors(yota_ctx* ctx);
#ifdef __CUDACC__
// Out-of-range entry `j1` (1-based, as the caller wrote it) of an index vector into an extent of n.
// Plain stores: every reporter writes the same flag; which offender's value survives is immaterial.
__device__ __forceinline__ void report_index_error(long long* err, long long j1, long long n) {
  if (err) {
    volatile long long* e = err;
    e[1] = j1;
    e[2] = n;
    e[0] = 1;
  }
}
#endif
int ctx_scratch(yota_ctx* ctx, size_t bytes, void** out);
// stream `s` waits for every in-flight reduction that overlaps [ptr, ptr + bytes)
int ctx_wait_pending(yota_ctx* ctx, const void* ptr, size_t bytes, cudaStream_t s);
// stream `s` waits for all in-flight reductions; the list is cleared
int ctx_join_comm(yota_ctx* ctx, cudaStream_t s);
// records a new in-flight reduction on the communication stream (event taken from the pool)
int ctx_add_pending(yota_ctx* ctx, const std::vector<std::pair<const char*, const char*>>& ranges);

// ---- elementwise region (ew_kernels.cu) ------------------------------------------------
// CLS_GATHER: a FULL operand read through a column index, element (r, c) = x[r, idx[c]]
// (getindex(x, :, idx) fused into its consumer region)
enum : int { CLS_FULL = 0, CLS_ROWVEC = 1, CLS_COLVEC = 2, CLS_SCALAR = 3, CLS_GATHER = 4 };
enum : int { OUT_STORE = 0, OUT_ROWSUM = 1, OUT_SCALARSUM = 2 };
constexpr int RG_MAX_IN = 12;
constexpr int RG_MAX_OUT = 12;
constexpr int RG_MAX_OPS = 64;
constexpr int RG_MAX_SLOTS = 28;
constexpr int RG_THREADS = 256;
constexpr uint8_t RG_ACC = 0xFF;   // operand: the previous micro-op's result
constexpr uint8_t RG_NONE = 0xFF;  // dst: keep only in the accumulator register

struct RegionIn {
  const float* ptr;
  int cls;
  int slot;
  const long long* idx;  // CLS_GATHER: 1-based Int64 column indices (out of range: reported, then clamped)
  long long gather_n;    // CLS_GATHER: number of columns of the gathered array
  long long* err;        // CLS_GATHER: where an out-of-range index is reported (yota_ctx::err_dev)
};
struct RegionOut {
  float* ptr;  // STORE: destination; ROWSUM / SCALARSUM: partials workspace
  int kind;
  int src;  // slot holding the value
  int acc;  // slot used as accumulator (sums)
  int _pad;
};
struct MicroOp {
  uint8_t op, dst, a, b;
  float imm;
};
struct RegionParams {
  long long R, C;
  int TRG, TCC, rowblocks, colchunks;
  long long cols_per_chunk;
  int n_in, n_out, n_ops, n_slots;
  RegionIn in[RG_MAX_IN];
  RegionOut out[RG_MAX_OUT];
  MicroOp ops[RG_MAX_OPS];
};

struct RegionLaunch {
  RegionParams p;
  int V;        // 4 or 1
  int grid;
  size_t smem;
  int static_id;  // >= 0: precompiled chain kernel
};

// geometry from (R, C, V): fills TRG, TCC, rowblocks, colchunks, cols_per_chunk, grid
void region_geometry(RegionLaunch& L, int sm_count, int ctas_per_sm = 8);
int launch_region(const RegionLaunch& L, cudaStream_t s);
// out[r] = scale * sum_p partial[p*R + r], p ascending (deterministic)
int launch_rowsum_finish(float* out, const float* partial, long long R, int P, float scale, cudaStream_t s);
// out[0] = scale * tree(partial[0..n))
int launch_scalar_finish(float* out, const float* partial, int n, float scale, cudaStream_t s);
// out[c] = scale * sum_r x[r + c*R]
int launch_colsum(float* out, const float* x, long long R, long long C, float scale, cudaStream_t s);

// generic N-d fallbacks (any broadcast pattern / any reduced-dims mask)
struct GenericEw {
  int opcode;
  int nd;
  long long dims[4];        // output dims
  long long sa[4], sb[4];   // element strides of the operands (0 = broadcast)
  float imm;
};
int launch_generic_ew(const GenericEw& g, float* out, const float* a, const float* b, cudaStream_t s);
struct GenericSum {
  int nd;
  long long dims[4];
  unsigned mask;
  float scale;
};
int launch_generic_sum(const GenericSum& g, float* out, const float* x, cudaStream_t s);
int launch_fill_const(float* out, float v, cudaStream_t s);

// static (precompiled) chain registry
int static_chain_lookup(const RegionParams& p, int V);
const char* static_chain_name(int id);
std::string region_signature(const RegionParams& p, int V);

// ---- GEMM (gemm_ffma.cu / gemm_tc.cu) ---------------------------------------------------
struct GemmArgs {
  int mode;  // 0 fp32 ffma, 1 tf32 tcgen05, 2 bf16 tcgen05, 3 compensated 3xTF32 tcgen05
  int transA, transB;
  long long M, N, K;
  const void* A;  // Float32 (modes 0, 1) or BF16 shadow (mode 2)
  long long lda;
  const void* B;
  long long ldb;
  float* C;
  long long ldc;
  int accumulate;
  // fused-epilogue GEMMs only: second, transposed store of chain output `tsh_out` (gemm_tc.cu)
  float* tsh = nullptr;
  long long tsh_ld = 0;
  int tsh_out = -1;
  // ... and BF16 copy of chain output `bsh_out` (BF16 mode: the operand shadow of later GEMMs)
  void* bsh = nullptr;
  int bsh_out = -1;
  // fused epilogue with a row-sum sink: final output, scale and arrival counters of the in-kernel
  // finisher (nullptr: the partials are reduced by a separate rowsum_finish launch)
  float* rs_out = nullptr;
  float rs_scale = 1.f;
  unsigned* rs_cnt = nullptr;
  // mode 3: Float32 shadows  x - tf32(x)  of both operands (same dims / leading dimensions)
  const void* A_lo = nullptr;
  const void* B_lo = nullptr;
  // 1-CTA tcgen05 kernel: launch with programmatic stream serialisation (the grid is scheduled
  // while the kernel before it in the stream drains; it waits for that kernel before it reads B,
  // the epilogue operands or writes anything).  early_a: A is not written by that kernel, so its
  // first tiles may be requested before the wait.
  int pdl = 0;
  int early_a = 0;
  // few output tiles, long K (the batched dW of an unrolled recurrence): the 2-CTA kernel cuts K
  // into `ksplit` slices (gemm_tc2_ksplit) whose partial tiles go to splitk_ws (ksplit * M * N
  // floats) and are summed in slice order by a second launch.  0 / nullptr: no split.
  int ksplit = 0;
  void* splitk_ws = nullptr;
};
// a whole recurrence of skinny GEMM + epilogue steps as one persistent launch (gemm_tc.cu); returns 1
// when the steps have to be launched one by one
int launch_gemm_chain(yota_ctx* ctx, const GemmArgs& g0, int static_id, const RegionParams& rp0, int z_in, int nsteps,
                      long long b_delta, long long e_dcol, unsigned* flags, cudaStream_t s);
// K slices the 2-CTA kernel would use for this GEMM (0: no split); sizes the workspace at plan time
int gemm_tc2_ksplit(const GemmArgs& g, int sm_count);
int launch_gemm_ffma(const GemmArgs& g, cudaStream_t s);
int launch_gemm_tc(yota_ctx* ctx, const GemmArgs& g, cudaStream_t s);  // returns YOTA_E_UNSUPPORTED_OP if shape not eligible
bool gemm_tc_eligible(const GemmArgs& g);
bool gemm_tc2_eligible(const GemmArgs& g, int sm_count);
bool gemm_epilogue_chain_ok(int static_id, int transA, int transB, int z_in);
// ... as an epilogue of the 1-CTA kernel (skinny outputs: TF32, 128 x 32 tiles, store-only chains)
bool gemm_epilogue1_chain_ok(const GemmArgs& g, int sm_count, int static_id, int z_in);
int launch_gemm_tc_epilogue(yota_ctx* ctx, const GemmArgs& g, int static_id, const RegionParams& rp, int z_in,
                            cudaStream_t s);
int launch_cast_bf16(void* dst, const float* src, long long n, cudaStream_t s);
int launch_split_tf32(float* lo, const float* src, long long n, cudaStream_t s);

// ---- indexing (index_kernels.cu) --------------------------------------------------------
struct IndexSpec {
  int n;                 // number of indices
  int kind[4];
  long long a[4], b[4];
  const int64_t* vec[4];  // device pointer for IDX_VEC
  long long vlen[4];
  int nd_x;
  long long xdims[4];     // dims of x as indexed (flattened when n == 1)
};
int launch_getindex(const IndexSpec& sp, float* y, const float* x, long long n_out, cudaStream_t s,
                    long long* err = nullptr);
size_t ungetindex_workspace_bytes(const IndexSpec& sp);
int ungetindex_launch_count(const IndexSpec& sp);
// A scatter-add is a sort of the index vector (needs nothing else) followed by the ordered fold of
// dy: UNGET_SORT / UNGET_FOLD run the halves separately on the same workspace, UNGET_ALL both.
enum { UNGET_ALL = 0, UNGET_SORT = 1, UNGET_FOLD = 2 };
int launch_ungetindex(const IndexSpec& sp, float* dx, const float* dy, void* workspace, size_t ws_bytes,
                      int sm_count, cudaStream_t s, const float* scale = nullptr, long long* err = nullptr,
                      int phase = UNGET_ALL);
// acc[I...] = acc[I...] + (0.0f + dy) over the selection (no index vectors); other elements untouched
int launch_slice_accum(const IndexSpec& sp, float* acc, const float* dy, bool scalar_dy, cudaStream_t s);
size_t scatter_cols_workspace_bytes(long long n_dst, long long n_src);
// scale != nullptr: folds scale[0] * dy[:, k] (one rounding per product, as if the product had
// been materialised first): the seed-scaled source of a scatter-add pullback read in place
int launch_scatter_add_cols(float* dx, long long rows, long long n_dst, const float* dy, const int64_t* idx,
                            long long n_src, void* workspace, size_t ws_bytes, cudaStream_t s,
                            const float* scale = nullptr, long long* err = nullptr, long long dst_lo = 0,
                            long long n_total = -1, int phase = UNGET_ALL);
int launch_gather_cols(float* y, const float* x, long long rows, long long n_dst, const int64_t* idx,
                       long long n_src, cudaStream_t s, long long* err = nullptr);

// ---- misc kernels (misc_kernels.cu) -----------------------------------------------------
int launch_logsoftmax(float* y, const float* x, long long R, long long C, cudaStream_t s);
int launch_transpose(float* y, const float* x, long long R, long long C, cudaStream_t s);
int launch_copy_strided(float* dst, const float* src, int nd, const long long* dims, const long long* dst_stride,
                        const long long* src_stride, cudaStream_t s);
int launch_axpy(float* x, const float* g, long long n, float lr, cudaStream_t s);
int launch_adam(float* x, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                const int* t_dev, int t_host, cudaStream_t s);
int launch_tick(int* t, cudaStream_t s);

// ---- NCCL (comm.cu) ---------------------------------------------------------------------
int comm_reserved_sms(yota_ctx* ctx);
// all-reduce (sum) of a bucket of buffers: the NVSwitch multimem kernel for buffers in symmetric
// memory, one grouped ncclAllReduce launch for the others
int comm_allreduce_bucket(yota_ctx* ctx, float* const* bufs, const long long* numel, int n, cudaStream_t s);

// ---- NVSwitch all-reduce over symmetric memory (symm_allreduce.cu) ------------------------
bool symm_available(yota_ctx* ctx);
int symm_alloc(yota_ctx* ctx, size_t bytes, void** out);
int symm_free(yota_ctx* ctx, void* ptr);
void symm_shutdown(yota_ctx* ctx);
bool symm_contains(yota_ctx* ctx, const void* ptr, size_t bytes);
int symm_ctas(yota_ctx* ctx);
bool symm_multimem(yota_ctx* ctx);
int symm_allreduce(yota_ctx* ctx, float* const* bufs, const long long* numel, int n, cudaStream_t s);

}  // namespace yota
