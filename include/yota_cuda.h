/* libyota_cuda.so -- C ABI of the B200-native execution backend for Yota.jl's gradient tape.
 *
 * This header is the drop-in boundary.  It replaces, for device arrays,
 * what the reference does at
 *   - src/grad.jl:293-307  grad_compile(tape): tape -> straight-line function   (yota_program_build)
 *   - src/grad.jl:353-355  invokelatest(gf, f, args...; seed)                   (yota_program_run)
 *   - src/grad.jl:147,173  rrule(cfg, f, args...) / pb(dy) arithmetic           (the op vocabulary below)
 *   - src/grad.jl:91       broadcast(+, dx, old_dx) tape gradient accumulation  (YOTA_EW_ADD, fused)
 *   - src/helpers.jl:7-11  ungetindex! -> NNlib.scatter!(+, dx, dy, idx)        (YOTA_OP_UNGETINDEX)
 *   - src/update.jl:42-49  update!(x, gx) = x .- gx                             (yota_update_sgd)
 * Plain pointers and sizes only; no C++/torch types.  All arrays are Julia arrays:
 * column-major, contiguous, Float32 data / Int64 (1-based) indices.
 *
 * Ownership: buffers bound as program inputs/outputs are caller-owned (yota_alloc / any
 * CUDA device pointer).  Every intermediate, workspace and saved activation is owned by
 * the program, sized at build time and reused by every run.  The library never keeps a
 * host pointer after a call returns.
 * Errors: every entry point returns 0 or a negative yota_status; the message is in the
 * thread-local yota_last_error().  There is NO CPU fallback: an op the backend cannot
 * lower fails yota_program_build with YOTA_E_UNSUPPORTED_OP (the analogue of
 * "No derivative rule found for op ..." at src/grad.jl:179-180).
 * Threading: one yota_ctx per device; calls on a ctx must be serialised by the caller.
 */
#ifndef YOTA_CUDA_H
#define YOTA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOTA_ABI_VERSION 1
#define YOTA_MAX_DIMS 4
#define YOTA_MAX_OP_INPUTS 6
#define YOTA_N_IATTR 16

typedef enum yota_status {
  YOTA_OK = 0,
  YOTA_E_CUDA = -1,
  YOTA_E_NCCL = -2,
  YOTA_E_SHAPE = -3,
  YOTA_E_UNSUPPORTED_OP = -4,
  YOTA_E_NOMEM = -5,
  YOTA_E_ARG = -6
} yota_status;

typedef enum yota_dtype { YOTA_F32 = 0, YOTA_I64 = 1 } yota_dtype;

/* Where a tensor of the op program lives. */
typedef enum yota_tensor_kind {
  YOTA_T_INPUT = 0,  /* bound at run time: in_ptrs[slot]                                   */
  YOTA_T_TEMP = 1,   /* program-owned intermediate (may never be materialised when fused)  */
  YOTA_T_OUTPUT = 2, /* bound at run time: out_ptrs[slot]                                  */
  YOTA_T_CONST = 3,  /* 0-dim Float32 constant, value = cval                               */
  YOTA_T_SEED = 4    /* 0-dim Float32 seed, value passed to yota_program_run (grad.jl:297) */
} yota_tensor_kind;

typedef struct yota_tensor_desc {
  int32_t dtype;               /* yota_dtype                                  */
  int32_t ndim;                /* 0..YOTA_MAX_DIMS (0 = scalar)               */
  int64_t dims[YOTA_MAX_DIMS]; /* Julia size(x): dims[0] is the fastest axis  */
  int32_t kind;                /* yota_tensor_kind                            */
  int32_t slot;                /* index into in_ptrs / out_ptrs               */
  double cval;                 /* YOTA_T_CONST value                          */
} yota_tensor_desc;

/* Op vocabulary: what the rrule(...) and pb(dy) statements of a compiled Yota tape
 *  bottom out in.  Elementwise ops follow Julia broadcast rules:
 * dims aligned from dim 1, missing trailing dims are 1, size-1 dims expand; the output
 * tensor's dims are the broadcast shape (YOTA_EW_COPY may expand to any larger shape:
 * this is the broadcast-back of the sum/mean pullback). */
typedef enum yota_opcode {
  YOTA_EW_COPY = 1,
  YOTA_EW_ADD = 2,
  YOTA_EW_SUB = 3,
  YOTA_EW_MUL = 4,
  YOTA_EW_DIV = 5,
  YOTA_EW_NEG = 6,
  YOTA_EW_TANH = 7,
  YOTA_EW_EXP = 8,
  YOTA_EW_LOG = 9,
  YOTA_EW_SIN = 10,
  YOTA_EW_COS = 11,
  YOTA_EW_SQRT = 12,
  YOTA_EW_RELU = 13,
  YOTA_EW_SIGMOID = 14,
  YOTA_EW_SQUARE = 15, /* x*x (abs2, literal_pow ^2)          */
  YOTA_EW_GT0 = 16,    /* x > 0 ? 1 : 0                       */
  YOTA_EW_ADDC = 17,   /* x + fattr[0]                        */
  YOTA_EW_MULC = 18,   /* x * fattr[0]                        */
  YOTA_EW_RSUBC = 19,  /* fattr[0] - x                        */
  YOTA_EW_RDIVC = 20,  /* fattr[0] / x                        */
  YOTA_EW_POWC = 21,   /* x ^ fattr[0]                        */
  YOTA_EW__LAST = 21,

  YOTA_OP_SUM = 64,       /* iattr[0] = bitmask of reduced dims (bit d <-> Julia dim d+1; all
                             dims + out ndim 0 = sum(x)); fattr[0] = post scale (1/n: mean) */
  YOTA_OP_MATMUL = 65,    /* out = op(A)*op(B); iattr[0] = transA, iattr[1] = transB;
                             vectors are n x 1 columns                                      */
  YOTA_OP_GETINDEX = 66,  /* out = x[I...]; see index spec below                            */
  YOTA_OP_UNGETINDEX = 67,/* out = zero(x); out[I...] += dy sequentially in column-major
                             source order, per destination a left fold from +0.0f           */
  YOTA_OP_LOGSOFTMAX = 68,/* NNlib.logsoftmax(x; dims=1)                                    */
  YOTA_OP_RESHAPE = 69,   /* same elements, new dims (alias)                                */
  YOTA_OP_TRANSPOSE = 70, /* 2-d transpose / vector -> 1 x n                                */
  YOTA_OP_CAT = 71,       /* cat(in...; dims=iattr[0])                                      */
  YOTA_OP_SLICE = 72      /* out = x[.., a:b, ..] along dim iattr[0] (1-based, iattr[1]=a)  */
} yota_opcode;

/* Index spec of GETINDEX / UNGETINDEX (Julia x[I...], 1-based):
 *   iattr[0] = n, the number of indices (1 on an N-d array = linear indexing)
 *   for index d (0-based): iattr[1+3d] = kind, iattr[2+3d] = a, iattr[3+3d] = b
 * GETINDEX inputs:   in[0] = x,  in[1..] = the index-vector tensors (Int64) in index order
 * UNGETINDEX inputs: in[0] = dy, in[1..] = the index-vector tensors; out dims = size(x)   */
typedef enum yota_index_kind {
  YOTA_IDX_COLON = 0, /* :                                  */
  YOTA_IDX_INT = 1,   /* a  (drops the dim)                 */
  YOTA_IDX_RANGE = 2, /* a:b inclusive                      */
  YOTA_IDX_VEC = 3    /* Int64 vector tensor, may repeat    */
} yota_index_kind;

typedef struct yota_op_desc {
  int32_t opcode; /* yota_opcode */
  int32_t n_in;
  int32_t in[YOTA_MAX_OP_INPUTS]; /* tensor ids */
  int32_t out;                    /* tensor id  */
  int32_t _pad;
  int64_t iattr[YOTA_N_IATTR];
  double fattr[2];
} yota_op_desc;

/* yota_program_build flags */
#define YOTA_FLAG_NO_FUSE 0x1u    /* one kernel per op (debug / A-B parity against fused)       */
#define YOTA_FLAG_NO_GRAPH 0x2u   /* launch step by step instead of replaying a CUDA graph      */
#define YOTA_FLAG_GEMM_FP32 0x0u  /* strict-FP32 FFMA contractions (default; rtol 1e-5 gate)    */
#define YOTA_FLAG_GEMM_TF32 0x10u /* tcgen05 kind::tf32 tensor-core contractions, FP32 accum    */
#define YOTA_FLAG_GEMM_BF16 0x20u /* tcgen05 kind::f16 (BF16 operands), FP32 accum              */
#define YOTA_FLAG_NO_STATIC 0x40u /* never pick a precompiled chain kernel (interpreter only)   */
#define YOTA_FLAG_GEMM_TF32X3 0x80u /* compensated tensor-core contractions: operands split into
                                     tf32 hi + lo, three tcgen05 MMAs per k-step (hi*hi + hi*lo +
                                     lo*hi) into one FP32 accumulator; meets the rtol 1e-5 gate   */

typedef struct yota_ctx yota_ctx;
typedef struct yota_program yota_program;

/* ---- library / context ---------------------------------------------------------------- */
int yota_abi_version(void);
const char* yota_last_error(void);
/* Creates the context on `device` (cudaSetDevice), one compute and one comm stream. */
int yota_init(int device, yota_ctx** out);
int yota_shutdown(yota_ctx* ctx);
int yota_device_info(yota_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- device memory (caller-owned buffers; the device-array type of the glue) ---------- */
int yota_alloc(yota_ctx* ctx, size_t bytes, void** dptr);
int yota_free(yota_ctx* ctx, void* dptr);
int yota_h2d(yota_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int yota_d2h(yota_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* pinned host staging (for overlapped uploads of each step's batch) */
int yota_host_alloc(yota_ctx* ctx, size_t bytes, void** hptr);
int yota_host_free(yota_ctx* ctx, void* hptr);
/* Asynchronous upload on the context's COPY stream (overlaps yota_program_run on the compute
 * stream).  Order the two streams with yota_stream_fence. */
int yota_h2d_async(yota_ctx* ctx, void* dst_dev, const void* src_pinned, size_t bytes);
/* Device -> pinned host copy on the compute stream, ordered after the runs issued before it; the
 * value is valid after yota_sync (or any later synchronous call).  Reads a step's loss without
 * stalling the host before it has enqueued the next step. */
int yota_d2h_async(yota_ctx* ctx, void* dst_pinned, const void* src_dev, size_t bytes);
/* Make stream `to` wait for everything enqueued so far on stream `from`
 * (0 = compute stream, 1 = copy stream). */
int yota_stream_fence(yota_ctx* ctx, int from, int to);
int yota_memset(yota_ctx* ctx, void* dptr, int byte, size_t bytes);
int yota_sync(yota_ctx* ctx);

/* ---- program: the compiled gradient tape ----------------------------------------------- */
/* Plans fusion, memory and launches for the op list.  No device work happens here, so it
 * also runs on a host without a GPU (used by the CPU-side tests of the planner). */
int yota_program_build(yota_ctx* ctx, const yota_op_desc* ops, int64_t n_ops,
                       const yota_tensor_desc* tensors, int64_t n_tensors, uint32_t flags,
                       yota_program** out);
/* One forward + pullback sweep.  in_ptrs/out_ptrs are device pointers indexed by
 * yota_tensor_desc.slot.  Asynchronous on the context's compute stream: read results
 * with yota_d2h (which orders after it) or call yota_sync. */
int yota_program_run(yota_program* p, void* const* in_ptrs, int64_t n_in, void* const* out_ptrs,
                     int64_t n_out, float seed);
int yota_program_destroy(yota_program* p);
/* Human-readable plan (steps, fused regions, chosen kernels, workspace bytes).  Returns the
 * number of bytes needed (excluding NUL); writes at most cap bytes. */
int64_t yota_program_describe(yota_program* p, char* buf, int64_t cap);
/* The op list as the planner sees it after its rewrite passes (unrolled-loop batching ...), one
 * tensor / op per text line (format: csrc/program.cu).  Diagnostic: lets a test execute the
 * rewritten program with an independent interpreter.  Same return convention as describe. */
int64_t yota_program_dump_ops(yota_program* p, char* buf, int64_t cap);
/* Counters: n_launches = kernels launched per run; workspace_bytes = program-owned memory. */
int yota_program_stats(yota_program* p, int64_t* n_steps, int64_t* n_launches,
                       int64_t* workspace_bytes, int64_t* n_fused_regions,
                       int64_t* n_static_kernels);
/* Device time of the last n runs' steps is not tracked here; use yota_event_* below. */

/* ---- timing on the compute stream (bench.py: CUDA events on the launching stream) ------ */
int yota_event_create(yota_ctx* ctx, void** ev);
int yota_event_record(yota_ctx* ctx, void* ev);
int yota_event_elapsed_ms(yota_ctx* ctx, void* ev_start, void* ev_stop, float* ms);
int yota_event_destroy(yota_ctx* ctx, void* ev);
/* Per-step device time of one run, measured with events around every step (serialises the
 * steps; diagnostic only).  names: '\n'-joined step labels. */
int yota_program_profile(yota_program* p, void* const* in_ptrs, int64_t n_in,
                         void* const* out_ptrs, int64_t n_out, float seed, float* step_ms,
                         int64_t cap_steps, char* names, int64_t cap_names);
/* Timeline of one eager run after `reps - 1` untimed ones, both streams: for every launched
 * step (compute stream) and every all-reduce bucket (communication stream) three floats --
 * device start, device end (ms since the run's first event) and host enqueue time (ms) -- and
 * a last record "(whole run)".  Shows how far the reductions overlap the pullback sweep.
 * Returns the number of records; names: '\n'-joined labels.  Diagnostic only. */
int yota_program_timeline(yota_program* p, void* const* in_ptrs, int64_t n_in,
                          void* const* out_ptrs, int64_t n_out, float seed, int reps,
                          float* rec_ms, int64_t cap_recs, char* names, int64_t cap_names);

/* ---- standalone operators (same kernels the program uses; unit-test / glue entry) ------ */
/* C[m x n] = op(A) op(B) (+ C if accumulate), column-major, lda/ldb/ldc in elements.
 * mode: 0 = strict FP32 FFMA, 1 = tcgen05 TF32, 2 = tcgen05 BF16, 3 = tcgen05 3xTF32 (compensated). */
int yota_gemm(yota_ctx* ctx, int mode, int transA, int transB, int64_t m, int64_t n, int64_t k,
              const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
              int accumulate);
/* dx[rows x n_dst] = zero; dx[:, idx[k]] += dy[:, k] for k = 1..n_src in order (bit-exact
 * sequential fold; src/helpers.jl:7-11).  idx is 1-based Int64 on the device. */
int yota_scatter_add_cols(yota_ctx* ctx, float* dx, int64_t rows, int64_t n_dst, const float* dy,
                          const int64_t* idx, int64_t n_src);
/* Destination-sharded form for multi-GPU use: this rank owns the table
 * columns dst_lo <= j < dst_hi (0-based, half open) of the n_dst and computes only them,
 * dx_shard[rows x (dst_hi - dst_lo)], from ALL sources (dy, idx replicated or gathered by the
 * caller).  No exchange and no reduction: every destination's fold keeps its global source
 * order, so the concatenation of the shards is bit-identical to yota_scatter_add_cols. */
int yota_scatter_add_cols_sharded(yota_ctx* ctx, float* dx_shard, int64_t rows, int64_t n_dst, int64_t dst_lo,
                                  int64_t dst_hi, const float* dy, const int64_t* idx, int64_t n_src);
/* y[:, k] = x[:, idx[k]] */
int yota_gather_cols(yota_ctx* ctx, float* y, const float* x, int64_t rows, int64_t n_dst,
                     const int64_t* idx, int64_t n_src);
/* x .= x .- lr .* gx   (update!(x, gx, (x,g) -> x .- lr.*g); lr = 1 is the reference default) */
int yota_update_sgd(yota_ctx* ctx, float* x, const float* gx, int64_t n, float lr);
/* One Adam step on the device, as Optimisers.jl's Adam applies it (the optimiser of the reference's
 * examples: examples/flux/mlp.jl:82, utils.jl:5): m, v are the caller's first / second moment
 * arrays (zero before step 1), t >= 1 the step number (bias correction 1 - beta^t). */
int yota_update_adam(yota_ctx* ctx, float* x, const float* gx, float* m, float* v, int64_t n, float lr,
                     float beta1, float beta2, float eps, int t);

/* ---- multi-GPU: batch-sharded data parallelism, NCCL all-reduce of parameter grads ------ */
/* 128-byte ncclUniqueId created on rank 0 and shipped to the others by the caller. */
int yota_comm_unique_id(void* id128);
int yota_comm_init(yota_ctx* ctx, int rank, int nranks, const void* id128);
int yota_comm_destroy(yota_ctx* ctx);
/* Marks program outputs (by out slot) whose buffers are summed over ranks inside
 * yota_program_run, each bucket all-reduced on the comm stream as soon as the step that
 * last writes it has been issued (overlaps the rest of the pullback sweep). */
int yota_program_set_allreduce(yota_program* p, const int32_t* out_slots, int64_t n);
int yota_allreduce_f32(yota_ctx* ctx, float* buf, int64_t n);
/* update! folded into the run (src/update.jl:42-49, fn = (x, gx) -> x .- lr .* gx): after the
 * sweep -- and after the all-reduce of the gradient, if it is marked -- the array bound to input
 * slot in_slots[i] is updated in place from the gradient in output slot out_slots[i].  One eval
 * then is one training step; n = 0 clears the marks. */
int yota_program_set_update(yota_program* p, const int32_t* in_slots, const int32_t* out_slots,
                            int64_t n, float lr);
/* Same marks, Adam instead of SGD: the moments and the step counter are owned by the program
 * (device memory, zero-initialised here; the counter advances once per run, on the device, so a
 * replayed CUDA graph takes correct steps).  Calling it again resets the optimiser state. */
int yota_program_set_update_adam(yota_program* p, const int32_t* in_slots, const int32_t* out_slots,
                                 int64_t n, float lr, float beta1, float beta2, float eps);
/* Symmetric memory for buffers that are all-reduced: collective (every rank calls it with the
 * same size, in the same order, after yota_comm_init).  Buffers carved from it are reduced by
 * the library's own NVSwitch kernel (multimem.ld_reduce / multimem.st over NCCL symmetric
 * windows, csrc/symm_allreduce.cu) instead of ncclAllReduce.  Needs NCCL >= 2.28;
 * yota_comm_symm_available: 0 = unavailable, 1 = available (peer pointers), 2 = multicast team
 * active (known once the first allocation has created the device communicator). */
int yota_comm_symm_available(yota_ctx* ctx);
int yota_comm_symm_alloc(yota_ctx* ctx, size_t bytes, void** dptr);
int yota_comm_symm_free(yota_ctx* ctx, void* dptr);

#ifdef __cplusplus
}
#endif
#endif /* YOTA_CUDA_H */
