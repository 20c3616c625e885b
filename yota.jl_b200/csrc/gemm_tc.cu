// Tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with TMEM accumulators, TMA-fed
// shared-memory pipeline, warp-specialised persistent CTAs.  Hand-written inline PTX; no
// CUTLASS / cuBLAS.
//
// This is the fast path of the matmul rrule (ChainRules rrule(*, A, B): forward A*B (NN),
// dA = Ȳ*B' (NT), dB = A'*Ȳ (TN) -- dispatched at /root/reference/src/grad.jl:147, forced at
// src/grad.jl:235).  Operands stay Float32 in HBM and are read by
// the tensor core as TF32 (10-bit mantissa), accumulation is FP32 in TMEM: stated tolerance
// ~1e-3 relative (tests/test_gpu_gemm_tc.py); the strict-FP32 FFMA kernel (gemm_ffma.cu)
// remains the rtol-1e-5 variant.
//
// Layout: everything is Julia column-major.  D[M x N] (M contiguous) = op(A)[M x K] op(B)[K x N].
//   op(A) with M contiguous in memory (A not transposed)  -> "MN-major" A operand
//   op(A) with K contiguous (A transposed)                 -> "K-major"  A operand
//   op(B) with K contiguous (B not transposed)             -> "K-major"  B operand
//   op(B) with N contiguous (B transposed)                 -> "MN-major" B operand
// so NN / NT / TN / TT need no transposes or copies: TMA loads the tile as it lies in HBM and
// the smem matrix descriptor tells the tensor core which way it is laid out.
//   K-major  tile: rows of 32 K-elements (128 B), SWIZZLE_128B;     k-step of 8 = +32 B
//   MN-major tile: rows of 32 MN-elements (128 B), one row per k, SWIZZLE_128B with 32-byte
//                  atoms (the only MN-major mode of 32-bit operands); 32-wide MN chunks are
//                  4 KiB apart (LBO), k-groups of 4 rows 512 B apart (SBO); k-step of 8 = +1 KiB
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue
// (TMEM -> registers -> coalesced 128-byte global stores; lane = row m, so a warp writes 32
// consecutive floats of one column).  4-stage smem ring (48 KiB / stage at 128x256x32), two
// TMEM accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the mainloop
// of tile i+1.  Persistent: grid = min(#tiles, #SMs), static round-robin tile order with the
// M index fastest so concurrently running CTAs share B panels in L2.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "ew_chain.cuh"

namespace yota {

namespace tc {

constexpr int TM = 128;      // tile M (= UMMA M, cta_group::1)
constexpr int STAGES = 4;
// EB = operand element bytes: 4 (Float32 read as TF32) or 2 (BF16 shadows of the Float32 arrays).
// A tile row is always one 128-byte swizzle row: TK = 128/EB elements; one MMA consumes 32 bytes
// of K: UMMA_K = 32/EB.  MN-major tiles: 128-byte rows of 128/EB MN-elements, one row per k.
template <int EB>
struct El {
  static constexpr int TK = 128 / EB;
  static constexpr int UMMA_K = 32 / EB;
  static constexpr int CHUNK = 128 / EB;                 // MN elements per 128-byte row
  static constexpr int CHUNK_BYTES = TK * 128;            // one MN-major box: TK rows x 128 B
  static constexpr int MN_SBO = EB == 4 ? 512 : 1024;     // k-group stride: 4-row (32B-atom) / 8-row atoms
  static constexpr int MN_LAYOUT = EB == 4 ? 1 : 2;       // SWIZZLE_128B_BASE32B / SWIZZLE_128B
  static constexpr int MN_KSTEP = UMMA_K * 128;           // bytes per MMA k-step in an MN-major tile
};
constexpr int THREADS = 192;
constexpr int EPI_WARP0 = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- cluster-scope helpers (2-CTA pairs, multicast clusters, split-K clusters) ----------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// release at cluster scope: the data this thread wrote into the target CTA's shared memory is
// visible to whoever acquires the barrier there
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAITC_DONE;\n\t"
      "bra WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(cta));
  return ra;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with programmatic stream serialisation may start
// while its predecessor in the stream still runs.  griddep_wait() returns once the predecessor has
// completed and its writes are visible (at once when there is no such dependency); nothing the
// predecessor may have written is read, and nothing is written to global memory, before it.
// griddep_launch() lets the NEXT kernel be scheduled early; it is only issued after this kernel's
// own wait, so that a successor's pre-wait work can never overlap anything older than this kernel.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane (row) into registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// ... 16 columns (the compensated mode's drain keeps 128 running sums per thread: a 32-column piece
// next to them does not fit in the 168 registers a thread of a 10-warp CTA gets)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor")
//  [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//  [46,48) version = 1       | [61,64) swizzle: 2 = 128B, 1 = 128B with 32-byte atoms
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Params {
  long long M, N, K;
  float* C;
  long long ldc;
  int accumulate;
  int num_m, num_n;
  int spread;  // 1: lane c of the producer warp issues TMA box c; 0: lane 0 issues all boxes
  // Compensated (3xTF32) mode only: k-blocks per accumulation chunk.  The tensor core adds every
  // MMA's partial sum into the FP32 accumulator with truncation, a bias that grows with the number
  // of accumulations (measured: 1.7e-4 of max|g| at K = 8192 x 3 products).  The accumulator is
  // therefore handed to the epilogue warps every `kchunk` k-blocks and added to C in the L2 with
  // round-to-nearest (red.global.add.f32; the tile stays in L2 between chunks), so the truncation
  // bias is bounded by the chunk length instead of K.
  int kchunk;
  // fused epilogue with a row-sum sink (db): the CTA that delivers the LAST partial of a 128-row
  // block also reduces the partials (same fixed order as rowsum_finish_kernel) and writes the
  // result -- no separate finisher launch on the critical path.  rs_cnt: one self-resetting
  // arrival counter per 128-row block.
  float* rs_out;
  float rs_scale;
  unsigned* rs_cnt;
  // 1-CTA kernel: operand A is not written by the kernel launched just before this one, so the A
  // tiles of the first ring pass may be requested before the grid dependency resolves
  int early_a;
  // 2-CTA kernel, plain stores only: the K range is cut into `ksplit` slices that are scheduled like
  // extra tiles (a 1024 x 1024 x 16384 contraction has 16 output tiles for 74 CTA pairs); slice i
  // stores its partial tile into part + i * M * N and splitk_reduce_kernel adds the slices in order.
  int ksplit;
  float* part;
  // fused epilogue: store output `tsh_out` of the chain a second time, transposed (element (m, n)
  // at tsh[n + m * tsh_ld]): the K-major operand of the dW GEMM that would otherwise read this
  // tensor MN-major (see plan_transposed_shadows in program.cu)
  float* tsh;
  long long tsh_ld;
  int tsh_out;
  // fused epilogue, BF16 mode: store output `bsh_out` a second time as BF16 (same layout): the
  // operand shadow the next GEMMs read, written here instead of by a cast kernel
  __nv_bfloat16* bsh;
  int bsh_out;
};

// ---- fused epilogues -------------------------------------------------------------------------
// The consumer region of a GEMM output (bias [+ tanh] after W*h; tanh'(h) * delta + db after
// W'*dz) runs on the accumulator tile instead of as a separate pass: the SAME ahead-of-time
// micro-program (ChainOf<ID>) the region kernel would execute, with the accumulator element
// standing in for the region's FULL operand `z_in`.  The epilogue's thread mapping (thread =
// row m, walking the tile's columns) is the region kernel's mapping, so ROWVEC operands (bias)
// are loaded once per tile and row sums (db) accumulate per thread in column order and leave
// as partial[tile_n][m] for the same fixed-order finisher.
// Out of line on purpose: the epilogue's hot loop keeps its register allocation (inlined, this tail
// pushed the tanh'/db epilogue kernel from 208 registers to 255 with spills and from 360 to 427 us).
// threadFenceReduction: publish this CTA's partials, count the arrival; the last of the num_n CTAs
// covering these 128 rows sums all partials in rowsum_finish_kernel's fixed order.
__device__ __noinline__ void finish_rowsum(const float* part, long long m, long long R, int num_n, float* out, float scale,
                                           unsigned* cnt, unsigned* flag) {
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
  const int et = (threadIdx.x - EPI_WARP0 * 32);
  if (et == 0) *flag = atomicAdd(cnt + (m >> 7), 1u) == (unsigned)(num_n - 1) ? 1u : 0u;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (*flag) {
    __threadfence();
    float sj[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float acc = 0.f;
      for (int q = j; q < num_n; q += 8) acc = __fadd_rn(acc, __ldcg(part + (long long)q * R + m));
      sj[j] = acc;
    }
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) t = __fadd_rn(t, sj[j]);
    out[m] = scale == 1.f ? t : __fmul_rn(t, scale);
    if (et == 0) cnt[m >> 7] = 0u;  // ready for the next run
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // the flag word is reused by the next tile
}
__device__ __noinline__ void prefetch_tile_l2(const float* line0, long long R, int lane) {
#pragma unroll
  for (int i = 0; i < 8; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(line0 + (long long)(lane + 32 * i) * R));
}

#define YCE(field, i) (ChainOf<ID>::value.field[decltype(i)::value])
template <int ID, int Z_IN>
struct Epilogue {
  static constexpr int N_IN = ChainOf<ID>::value.n_in, N_OPS = ChainOf<ID>::value.n_ops, N_OUT = ChainOf<ID>::value.n_out;
  static constexpr int NS = ChainOf<ID>::value.n_slots > 0 ? ChainOf<ID>::value.n_slots : 1;
  static constexpr int NO = N_OUT > 0 ? N_OUT : 1;
  static constexpr int first_rowsum() {
    for (int k = 0; k < N_OUT; k++)
      if (ChainOf<ID>::value.out_kind[k] == OUT_ROWSUM) return k;
    return -1;
  }
  static constexpr int RS_OUT = first_rowsum();  // the sink whose partials the kernel finishes itself

  __device__ __forceinline__ static void begin_tile(const RegionParams& rp, long long m, long long n0, long long R,
                                                    float (&base)[NS], float (&accs)[NO], bool prefetch) {
    // With ONE tile per CTA (small batches per GPU) nothing hides the latency of the epilogue's
    // operand loads -- beside an all-reduce that latency triples -- so the tile's other FULL
    // operands (h of the tanh' epilogue) are pulled into L2 while the mainloop runs.  With several
    // tiles per CTA the next mainloop hides the epilogue and a prefetch would only be evicted by
    // the operand stream before it is used (measured: +180 MB of DRAM reads, 360 -> 426 us).
    static_for<N_IN>([&](auto k) {
      constexpr int K = decltype(k)::value, cls = YCE(in_cls, k);
      if constexpr (cls == CLS_FULL && K != Z_IN) if (prefetch) {
        const int lane = threadIdx.x & 31;
        prefetch_tile_l2(rp.in[K].ptr + n0 * R + (m - lane), R, lane);
      }
    });
    static_for<NS>([&](auto s) { base[decltype(s)::value] = 0.f; });
    static_for<NO>([&](auto k) { accs[decltype(k)::value] = 0.f; });
    static_for<N_IN>([&](auto k) {
      constexpr int cls = YCE(in_cls, k), slot = YCE(in_slot, k);
      if constexpr (cls == CLS_ROWVEC) base[slot] = __ldg(rp.in[decltype(k)::value].ptr + m);
      if constexpr (cls == CLS_SCALAR) base[slot] = __ldg(rp.in[decltype(k)::value].ptr);
    });
  }
  // one chunk of 32 consecutive columns: accumulator values z[j] at (m, n0 + j).  The other FULL
  // operands of the chunk are loaded up front (32 independent loads in flight per operand) --
  // left per element, each load sits between the previous element's store and its own use and
  // the epilogue runs at one DRAM latency per element.
  static constexpr int NI = N_IN > 0 ? N_IN : 1;
  __device__ __forceinline__ static void chunk32(const RegionParams& rp, const uint32_t (&z)[32], long long m,
                                                 long long n0, long long R, const float (&base)[NS], float (&accs)[NO],
                                                 float* tsh, long long tsh_ld, int tsh_out, float4* stage,
                                                 __nv_bfloat16* bsh, int bsh_out) {
    float pre[NI][32];
    load_operands(rp, m, n0, R, pre);
    chunk32_with(rp, z, pre, m, n0, R, base, accs, tsh, tsh_ld, tsh_out, stage, bsh, bsh_out);
  }
  // the chunk's other operands (everything but the accumulator): independent of the GEMM, so a caller
  // may request them before it waits for the accumulator (the persistent recurrence kernel does)
  __device__ __forceinline__ static void load_operands(const RegionParams& rp, long long m, long long n0, long long R,
                                                       float (&pre)[NI][32]) {
    static_for<N_IN>([&](auto k) {
      constexpr int K = decltype(k)::value, cls = YCE(in_cls, k);
      if constexpr (cls == CLS_FULL && K != Z_IN) {
        const float* src = rp.in[K].ptr + n0 * R + m;
#pragma unroll
        for (int j = 0; j < 32; j++) pre[K][j] = __ldg(src + (long long)j * R);
      }
      if constexpr (cls == CLS_COLVEC) {
        // lanes read consecutive entries once and broadcast them with shuffles
        const float mine = __ldg(rp.in[K].ptr + n0 + (threadIdx.x & 31));
#pragma unroll
        for (int j = 0; j < 32; j++) pre[K][j] = __shfl_sync(0xffffffffu, mine, j);
      }
    });
  }
  __device__ __forceinline__ static void chunk32_with(const RegionParams& rp, const uint32_t (&z)[32],
                                                      const float (&pre)[NI][32], long long m, long long n0, long long R,
                                                      const float (&base)[NS], float (&accs)[NO], float* tsh,
                                                      long long tsh_ld, int tsh_out, float4* stage, __nv_bfloat16* bsh,
                                                      int bsh_out) {
    const int lane = threadIdx.x & 31;
    float tv[4];
#pragma unroll
    for (int j = 0; j < 32; j++) {
      float v[NS];
      static_for<NS>([&](auto s) { v[decltype(s)::value] = base[decltype(s)::value]; });
      const long long off = (n0 + j) * R + m;
      static_for<N_IN>([&](auto k) {
        constexpr int K = decltype(k)::value, cls = YCE(in_cls, k), slot = YCE(in_slot, k);
        if constexpr (cls == CLS_FULL) v[slot] = (K == Z_IN) ? __uint_as_float(z[j]) : pre[K][j];
        if constexpr (cls == CLS_COLVEC) v[slot] = pre[K][j];
      });
      float acc = 0.f;
      static_for<N_OPS>([&](auto i) {
        constexpr int OP = YCE(op, i), D = YCE(dst, i), A = YCE(a, i), B = YCE(b, i);
        float a, b;
        if constexpr (A == RG_ACC) a = acc; else a = v[A];
        if constexpr (B == RG_ACC) b = acc; else b = v[B];
        acc = ew_v<OP>(a, b, rp.ops[decltype(i)::value].imm);
        if constexpr (D != RG_NONE) v[D] = acc;
      });
      static_for<N_OUT>([&](auto k) {
        constexpr int kind = YCE(out_kind, k), src = YCE(out_src, k);
        if constexpr (kind == OUT_STORE) {
          rp.out[decltype(k)::value].ptr[off] = v[src];
          if (decltype(k)::value == tsh_out) tv[j & 3] = v[src];
          if (bsh && decltype(k)::value == bsh_out) bsh[off] = __float2bfloat16_rn(v[src]);
        } else
          accs[decltype(k)::value] = __fadd_rn(accs[decltype(k)::value], v[src]);
      });
      // this thread's 32 columns of row m are one 128-byte line of the transposed copy: staged in
      // shared memory (16-byte pieces XOR-swizzled by row: conflict-free both ways) ...
      if ((j & 3) == 3 && tsh) stage[lane * 8 + ((j >> 2) ^ (lane & 7))] = make_float4(tv[0], tv[1], tv[2], tv[3]);
    }
    if (tsh) {
      // ... and written out 4 whole lines per warp instruction (8 lanes x 16 bytes each) instead of
      // 32 quarter-sector stores to 32 different lines
      __syncwarp();
      float* base_row = tsh + (m - lane) * tsh_ld + n0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int rr = 4 * i + (lane >> 3), piece = lane & 7;
        reinterpret_cast<float4*>(base_row + (long long)rr * tsh_ld)[piece] = stage[rr * 8 + (piece ^ (rr & 7))];
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ static void end_tile(const RegionParams& rp, long long m, long long tile_n, long long R,
                                                  const float (&accs)[NO], const Params& p, unsigned* flag) {
    static_for<N_OUT>([&](auto k) {
      constexpr int kind = YCE(out_kind, k);
      if constexpr (kind == OUT_ROWSUM) rp.out[decltype(k)::value].ptr[tile_n * R + m] = accs[decltype(k)::value];
    });
    constexpr int RS = RS_OUT;
    if constexpr (RS >= 0) {
      if (p.rs_cnt) finish_rowsum(rp.out[RS].ptr, m, R, p.num_n, p.rs_out, p.rs_scale, p.rs_cnt, flag);
    }
  }

};
#undef YCE

// X3 (3xTF32, the compensated mode): a stage also holds the LOW parts of both tiles,
// lo = x - tf32(x) (planner-made Float32 shadows, see launch_split_tf32), and every k-step
// issues three MMAs into the same TMEM accumulator: hi*hi + hi*lo + lo*hi.  What is dropped is
// lo*lo and the truncation of lo itself, ~2^-21 relative per product -- the error level of an
// FP32 FMA chain, so this mode is gated at the reference's rtol 1e-5 (tests/gpu_util.py::gate)
// while running on the tensor cores at ~1/3 of the TF32 rate.
template <int TN, bool X3 = false>
struct Smem {
  static constexpr int A_BYTES = TM * 128;  // TM rows (or TM*EB/128 chunks of TK rows) of 128 bytes
  static constexpr int B_BYTES = TN * 128;
  static constexpr int HALF = A_BYTES + B_BYTES;  // the hi tiles; X3: the lo tiles follow
  static constexpr int STAGE_BYTES = X3 ? 2 * HALF : HALF;
  // narrow tiles (skinny outputs: a 1024 x 128 result is 8 CTAs at TN = 128, 32 at TN = 32) are
  // latency-bound streams of A: twice the stages keep the same bytes in flight
  static constexpr int NSTAGE = X3 ? (TN == 256 ? 2 : TN == 128 ? 3 : 4) : (TN <= 64 ? 2 * STAGES : STAGES);
  static constexpr int BAR_OFF = NSTAGE * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // + barriers + slack for 1 KiB alignment
};

// EPI >= 0 (skinny outputs of an unrolled recurrence, C4: 1024 x 128 per step): the consumer region's
// chain runs on the accumulator like in the 2-CTA kernel -- h = tanh.(zx .+ Wh*h .+ b) forward,
// dz = (Wh'*dz) .* (1 .- h.^2) backward -- stores only (no row-sum sinks, no shadows).
// KS > 1 (skinny outputs with a long K: the per-step GEMMs of a recurrence, 1024 x 128 x 1024):
// a cluster of KS CTAs shares one output tile, each streaming 1/KS of K -- a CTA that streams all
// of its A panel alone is bound by what one SM can pull from L2 (640 KB in ~7 us), four of them
// need a quarter of that.  The non-leader CTAs write their accumulators into the leader's shared
// memory (DSMEM), the leader adds them in rank order (deterministic) and runs the epilogue.
template <int TN, bool A_MN, bool B_MN, int EB, bool X3, int EPI, int ZIN, int KS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_al, const __grid_constant__ CUtensorMap map_bl, const Params p,
               const __grid_constant__ RegionParams rp) {
  static_assert(!X3 || EB == 4, "the compensated mode splits Float32 operands");
  using S = Smem<TN, X3>;
  using E = El<EB>;
  constexpr int TK = E::TK;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + S::NSTAGE;
  uint64_t* tfull = empty + S::NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  static_assert(!(KS > 1) || (!X3 && S::NSTAGE * 16 + 2 * 16 + 8 <= 176), "split-K: plain TF32/BF16, barriers fit");
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF + 176);   // leader: partial tiles of the other CTAs landed
  uint64_t* pempty = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF + 184);  // non-leaders: the leader has consumed them
  float* part = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);          // leader: [KS - 1][TN][128] partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m * p.num_n;
  const uint32_t kr = KS > 1 ? cluster_ctarank() : 0u;  // which K-slice of the tile this CTA computes
  const int tile0 = (int)blockIdx.x / KS, tile_step = (int)gridDim.x / KS;
  const int nkb = (int)(p.K / (128 / EB)) / KS;   // k-blocks of this CTA ...
  const int kbeg = (int)kr * nkb;                 // ... starting here

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::NSTAGE; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
    }
    if (KS > 1) {
      mbar_init(pfull, (KS - 1) * 128);
      mbar_init(pempty, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulator stages x TN columns (512 when TN = 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * TN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (KS > 1) cluster_sync_all();  // every CTA's barriers exist before anyone arrives remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer warp =====================
    // An MN-major operand needs one box per 128-byte-wide MN chunk (up to 4 + 8 boxes per
    // stage); issuing them from one thread costs more than the stage's MMA time, so lane c
    // issues box c: the whole stage is in flight after one issue slot.
    constexpr int NA = A_MN ? TM / E::CHUNK : 1;
    constexpr int NB = B_MN ? TN / E::CHUNK : 1;
    static_assert((X3 ? 2 : 1) * (NA + NB) <= 32, "one lane per TMA box");
    // issue the TMA boxes of operand A and / or B for k-block k0 into `stage`
    auto issue = [&](int stage, int m0, int n0, int k0, bool do_a, bool do_b) {
#pragma unroll
      for (int part = 0; part < (X3 ? 2 : 1); part++) {  // part 1: the lo tiles of the compensated mode
        unsigned char* sa = smem + stage * S::STAGE_BYTES + part * S::HALF;
        unsigned char* sb = sa + S::A_BYTES;
        const CUtensorMap* ma = part ? &map_al : &map_a;
        const CUtensorMap* mb = part ? &map_bl : &map_b;
        const int l0 = part * (NA + NB);
        if (!p.spread) {
          if (lane == 0) {
            if (do_a) {
              if (A_MN) {
#pragma unroll
                for (int c = 0; c < NA; c++) tma_load_2d(sa + c * E::CHUNK_BYTES, ma, &full[stage], m0 + E::CHUNK * c, k0);
              } else
                tma_load_2d(sa, ma, &full[stage], k0, m0);
            }
            if (do_b) {
              if (B_MN) {
#pragma unroll
                for (int c = 0; c < NB; c++) tma_load_2d(sb + c * E::CHUNK_BYTES, mb, &full[stage], n0 + E::CHUNK * c, k0);
              } else
                tma_load_2d(sb, mb, &full[stage], k0, n0);
            }
          }
        } else if (lane >= l0 && lane < l0 + NA) {
          if (do_a) {
            if (A_MN)
              tma_load_2d(sa + (lane - l0) * E::CHUNK_BYTES, ma, &full[stage], m0 + E::CHUNK * (lane - l0), k0);
            else
              tma_load_2d(sa, ma, &full[stage], k0, m0);
          }
        } else if (lane >= l0 + NA && lane < l0 + NA + NB) {
          if (do_b) {
            const int c = lane - l0 - NA;
            if (B_MN)
              tma_load_2d(sb + c * E::CHUNK_BYTES, mb, &full[stage], n0 + E::CHUNK * c, k0);
            else
              tma_load_2d(sb, mb, &full[stage], k0, n0);
          }
        }
      }
    };
    // Before the grid dependency resolves: the A tiles of the first ring pass (the ring is empty,
    // nothing to wait for), when A cannot have been written by the kernel still running.
    int pre = 0;
    if (p.early_a && tile0 < num_tiles) {
      pre = nkb < S::NSTAGE ? nkb : S::NSTAGE;
      const int m0 = (tile0 % p.num_m) * TM;
      for (int kb = 0; kb < pre; kb++) {
        if (lane == 0) mbar_expect_tx(&full[kb], S::STAGE_BYTES);
        __syncwarp();
        issue(kb, m0, 0, (kbeg + kb) * TK, true, false);
      }
    }
    griddep_wait();
    griddep_launch();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = tile0; t < num_tiles; t += tile_step) {
      const int m0 = (t % p.num_m) * TM, n0 = (t / p.num_m) * TN;
      for (int kb = 0; kb < nkb; kb++) {
        const bool half = t == tile0 && kb < pre;  // A is already in flight, the barrier already armed
        if (!half) {
          if (lane == 0) {  // one poller per warp; the other lanes are released by the warp barrier
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], S::STAGE_BYTES);
          }
          __syncwarp();
        }
        issue(stage, m0, n0, (kbeg + kb) * TK, !half, true);
        if (++stage == S::NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    griddep_wait();
    griddep_launch();
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32 (2) or BF16 (1), majors, N >> 3, M >> 4
      constexpr uint32_t FMT = EB == 4 ? 2u : 1u;
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int kc = X3 ? p.kchunk : nkb;
      for (int t = tile0; t < num_tiles; t += tile_step)
       for (int ch = 0; ch < (X3 ? (nkb + kc - 1) / kc : 1); ch++) {  // one accumulator per chunk of k-blocks (X3) / per tile
        const int kb0 = X3 ? ch * kc : 0;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN;
        const int kend = kb0 + kc < nkb ? kb0 + kc : nkb;
        for (int kb = kb0; kb < kend; kb++) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; k++) {  // 4 MMA k-steps of 32 bytes per 128-byte stage
            const uint64_t da = A_MN ? make_desc(sa + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                     : make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = B_MN ? make_desc(sb + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                     : make_desc(sb + k * 32, 16, 1024, 2);
            if (EB == 4)
              umma_tf32(d_tmem, da, db, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else
              umma_f16(d_tmem, da, db, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            if constexpr (X3) {
              const uint32_t sal = sa + S::HALF, sbl = sal + S::A_BYTES;
              const uint64_t dal = A_MN ? make_desc(sal + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                        : make_desc(sal + k * 32, 16, 1024, 2);
              const uint64_t dbl = B_MN ? make_desc(sbl + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                        : make_desc(sbl + k * 32, 16, 1024, 2);
              umma_tf32(d_tmem, da, dbl, idesc, 1u);
              umma_tf32(d_tmem, dal, db, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == S::NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    griddep_wait();  // epilogue operands and C itself may come from the kernel before this one
    griddep_launch();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t pphase = 0;  // split-K: phase of the partial-sum hand-over barriers
    const int kc = X3 ? p.kchunk : nkb;
    for (int t = tile0; t < num_tiles; t += tile_step)
     for (int ch = 0; ch < (X3 ? (nkb + kc - 1) / kc : 1); ch++) {
      const int kb0 = X3 ? ch * kc : 0;
      const long long m = (long long)(t % p.num_m) * TM + q * 32 + lane;
      const long long n0 = (long long)(t / p.num_m) * TN;
      // X3: every chunk after the first (and the first too when C += ...) is ADDED to C by the L2
      // (red.global.add.f32: round-to-nearest, no read on the SM side; one thread owns an element,
      // so its adds arrive in chunk order -- deterministic)
      const bool red = X3 && (p.accumulate || kb0 > 0);
      const bool add_old = !X3 && p.accumulate;
      if constexpr (KS > 1) {
        if (kr != 0) {
          // partial sums of this CTA's K-slice -> the leader's shared memory, [slice][column][row]
          mbar_wait(pempty, pphase ^ 1);  // the leader is done with the previous tile's partials
          mbar_wait(&tfull[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t dst0 = map_to_cta(part, 0) + (uint32_t)((((kr - 1) * TN) * 128 + (q * 32 + lane)) * 4);
#pragma unroll 1
          for (int c = 0; c < TN / 32; c++) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + (uint32_t)(acc * TN + c * 32) + ((uint32_t)(q * 32) << 16);
            tmem_ld32(taddr, v);
#pragma unroll
            for (int j = 0; j < 32; j++) st_cluster_f32(dst0 + (uint32_t)((c * 32 + j) * 128 * 4), __uint_as_float(v[j]));
          }
          tcgen05_fence_before();
          mbar_arrive(&tempty[acc]);
          mbar_arrive_remote_release(pfull, 0);
          pphase ^= 1;
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
          continue;
        }
      }
      using EP = Epilogue<(EPI >= 0 ? EPI : 0), ZIN>;
      float ebase[EP::NS], eaccs[EP::NO];
      if constexpr (EPI >= 0) EP::begin_tile(rp, m, n0, p.M, ebase, eaccs, false);
      mbar_wait(&tfull[acc], acc_phase);
      if constexpr (KS > 1) {
        mbar_wait_acquire_cluster(pfull, pphase);  // the other K-slices' partial sums have landed
        pphase ^= 1;
      }
      tcgen05_fence_after();
      float* crow = p.C + m + n0 * p.ldc;
#pragma unroll 1
      for (int c = 0; c < TN / 32; c++) {
        uint32_t v[32];
        // C += ...: the 32 old values are requested before the accumulator read, all in flight at
        // once (left inside the store loop they serialise: load -> add -> store per element)
        float old[32];
        if (EPI < 0 && add_old) {
          const float* cp0 = crow + (long long)(c * 32) * p.ldc;
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = __ldcg(cp0 + j * p.ldc);
        }
        const uint32_t taddr = tmem_base + (uint32_t)(acc * TN + c * 32) + ((uint32_t)(q * 32) << 16);
        tmem_ld32(taddr, v);
        if constexpr (KS > 1) {  // + the other K-slices, in rank order
#pragma unroll
          for (int r = 1; r < KS; r++) {
            const float* pr = part + ((r - 1) * TN + c * 32) * 128 + (q * 32 + lane);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), pr[j * 128]));
          }
        }
        float* cp = crow + (long long)(c * 32) * p.ldc;
        if constexpr (EPI >= 0) {
          EP::chunk32(rp, v, m, n0 + c * 32, p.M, ebase, eaccs, nullptr, 0, -1, nullptr, nullptr, -1);
        } else if (X3 && red) {
#pragma unroll
          for (int j = 0; j < 32; j++) atomicAdd(cp + j * p.ldc, __uint_as_float(v[j]));
        } else if (add_old) {
#pragma unroll
          for (int j = 0; j < 32; j++) cp[j * p.ldc] = __uint_as_float(v[j]) + old[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) cp[j * p.ldc] = __uint_as_float(v[j]);
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty[acc]);
      if constexpr (KS > 1) {  // the partial-sum buffer may be overwritten for the next tile
#pragma unroll
        for (int r = 1; r < KS; r++) mbar_arrive_remote(pempty, (uint32_t)r);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (KS > 1) cluster_sync_all();  // the leader's shared memory outlives the others' DSMEM stores
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TN) : "memory");
  }
}

// =============================================================================================
// Persistent recurrence: h_i = chain(A * h_{i-1}, operands_i) for i = 0 .. nsteps-1 in ONE launch.
// The per-step products of an unrolled RNN (1024 x 128 x 1024, ~8 us per launch of the split-K
// kernel above, most of it prologue, the A stream and the launch hand-over) repeat the same A
// (the recurrence weights) on the previous step's output.  Here a cluster of KS = 4 CTAs owns
// one 128 x 32 output tile for the whole recurrence: its K-slice of A is loaded into the smem
// ring ONCE and stays there (K / KS <= 8 k-blocks), per step only the 32 KiB B slice moves.
// Steps are ordered through global counters, one per output tile: each of the leader's 128 epilogue
// threads adds 1 (release) after it stored its part of h_i; a producer whose K-slice covers rows
// [r0, r1) of h_i polls the counters of the tiles holding those rows (acquire) until they show
// 128 (i + 1), then crosses into the async proxy (fence.proxy.async) before the TMA loads.  All
// clusters are co-resident (the host checks), nothing else on the device waits for this kernel
// while it spins, and a poll that never succeeds gives up after ~1 s (wait_step_flag).
// Operands advance per step by whole columns: B tile column b_col0 + i * b_dcol of one tensor map
// over the container of all h_i; the chain's FULL operands and stores by e_dcol columns (folded
// into the column argument of the epilogue -- ROWVEC / SCALAR operands do not move).
struct ChainParams {
  int nsteps;
  int b_col0, b_dcol;
  long long e_dcol;
  unsigned* flags;  // one counter per output tile, zeroed by the host before the launch
  long long* err;  // err[3] = 1: a counter did not arrive (reported by the next synchronising call)
  // diagnostics (YOTA_CHAIN_TRACE=1, tools/chain_trace.py): %globaltimer stamps of the cluster of tile 0,
  // [step][8]: 0 flags seen, 1 B requested, 2 products done (leader), 3 products done (rank 1),
  // 4 partial sums sent (rank 1), 5 partial sums in (leader), 6 stored, 7 published
  unsigned long long* trace;
};
__device__ __forceinline__ void chain_stamp(const ChainParams& cp, int tile, int i, int what) {
  if (cp.trace && tile == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    cp.trace[i * 8 + what] = t;
  }
}

// false: gave up (~1 s).  The results are garbage then, but nothing hangs: the caller stops waiting
// for the rest of the recurrence and the next synchronising call reports the failure.
__device__ __forceinline__ bool wait_step_flag(const unsigned* f, unsigned want, long long* err) {
  for (int spin = 0; spin < (1 << 20); spin++) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v >= want) return true;
    __nanosleep(20);
  }
  if (err) *(volatile long long*)(err + 3) = 1;
  return false;
}

template <bool A_MN, int EPI, int ZIN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_chain_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p,
                  const ChainParams cp, const __grid_constant__ RegionParams rp) {
  constexpr int TN = 32, KS = 4, EB = 4;
  using S = Smem<TN, false>;
  using E = El<EB>;
  constexpr int TK = E::TK;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + S::NSTAGE;
  uint64_t* tfull = empty + S::NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  static_assert(S::NSTAGE * 16 + 2 * 16 + 8 <= 176, "barriers fit");
  uint64_t* pfull = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF + 176);
  uint64_t* pempty = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF + 184);
  float* part = reinterpret_cast<float*>(smem + S::BAR_OFF + 256);  // leader: [KS - 1][TN][128] partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t kr = cluster_ctarank();
  const int tile = (int)blockIdx.x / KS;  // one tile per cluster for the whole recurrence
  const int tm = tile % p.num_m, tn = tile / p.num_m;
  const int nkb = (int)(p.K / TK) / KS;  // <= NSTAGE: k-block kb of this CTA lives in stage kb
  const int kbeg = (int)kr * nkb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::NSTAGE; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 128);
    }
    mbar_init(pfull, (KS - 1) * 128);
    mbar_init(pempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * TN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: lane kb owns k-block kb =====================
    // (one lane issuing all eight loads of a step took 0.77 us of the 7.4 us step; the two flags of
    // a step are polled by two lanes at once)
    constexpr int NA = A_MN ? TM / E::CHUNK : 1;
    const int m0 = tm * TM, n0 = tn * TN;
    // output tiles of the previous step that hold rows [kbeg * TK, (kbeg + nkb) * TK) of B
    const int t_lo = (kbeg * TK) / TM, t_hi = ((kbeg + nkb) * TK - 1) / TM;
    bool alive = true;
    for (int i = 0; i < cp.nsteps; i++) {
      if (i > 0) {
        for (int t = t_lo + lane; t <= t_hi; t += 32)
          if (alive) alive = wait_step_flag(cp.flags + tn * p.num_m + t, (unsigned)i * 128u, cp.err);
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores of other SMs -> TMA reads
      }
      if (kr == 0 && lane == 0) chain_stamp(cp, tile, i, 0);
      const int nb = cp.b_col0 + i * cp.b_dcol + n0;
      if (lane < nkb) {
        const int kb = lane;
        mbar_wait(&empty[kb], (uint32_t)((i & 1) ^ 1));
        unsigned char* sa = smem + kb * S::STAGE_BYTES;
        const int k0 = (kbeg + kb) * TK;
        if (i == 0) {  // A: once, resident for all steps
          mbar_expect_tx(&full[kb], S::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < NA; c++) tma_load_2d(sa + c * E::CHUNK_BYTES, &map_a, &full[kb], m0 + E::CHUNK * c, k0);
          } else
            tma_load_2d(sa, &map_a, &full[kb], k0, m0);
        } else
          mbar_expect_tx(&full[kb], S::B_BYTES);
        tma_load_2d(sa + S::A_BYTES, &map_b, &full[kb], k0, nb);
      }
      if (kr == 0 && lane == 0) chain_stamp(cp, tile, i, 1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((uint32_t)(TN >> 3) << 17) |
                             ((uint32_t)(TM >> 4) << 24);
      for (int i = 0; i < cp.nsteps; i++) {
        const int acc = i & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((i >> 1) & 1) ^ 1));
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&full[kb], (uint32_t)(i & 1));
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + kb * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t da = A_MN ? make_desc(sa + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                     : make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = make_desc(sb + k * 32, 16, 1024, 2);
            umma_tf32(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[kb]);  // B of this stage may be overwritten by the next step
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const long long m = (long long)tm * TM + q * 32 + lane;
    for (int i = 0; i < cp.nsteps; i++) {
      const int acc = i & 1;
      const uint32_t acc_par = (uint32_t)((i >> 1) & 1), ppar = (uint32_t)(i & 1);
      if (kr != 0) {
        mbar_wait(pempty, ppar ^ 1);
        mbar_wait(&tfull[acc], acc_par);
        tcgen05_fence_after();
        if (kr == 1 && threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 3);
        const uint32_t dst0 = map_to_cta(part, 0) + (uint32_t)((((kr - 1) * TN) * 128 + (q * 32 + lane)) * 4);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t)(acc * TN) + ((uint32_t)(q * 32) << 16);
        tmem_ld32(taddr, v);
#pragma unroll
        for (int j = 0; j < 32; j++) st_cluster_f32(dst0 + (uint32_t)(j * 128 * 4), __uint_as_float(v[j]));
        tcgen05_fence_before();
        mbar_arrive(&tempty[acc]);
        mbar_arrive_remote_release(pfull, 0);
        if (kr == 1 && threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 4);
        continue;
      }
      using EP = Epilogue<EPI, ZIN>;
      float ebase[EP::NS], eaccs[EP::NO];
      const long long n0 = (long long)tn * TN + (long long)i * cp.e_dcol;  // this step's columns of the containers
      EP::begin_tile(rp, m, n0, p.M, ebase, eaccs, false);
      float pre[EP::NI][32];
      EP::load_operands(rp, m, n0, p.M, pre);  // in flight while this step's products are still being computed
      mbar_wait(&tfull[acc], acc_par);
      if (threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 2);
      mbar_wait_acquire_cluster(pfull, ppar);
      tcgen05_fence_after();
      if (threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 5);
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (uint32_t)(acc * TN) + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, v);
#pragma unroll
      for (int r = 1; r < KS; r++) {  // + the other K-slices, in rank order
        const float* pr = part + ((r - 1) * TN) * 128 + (q * 32 + lane);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __float_as_uint(__fadd_rn(__uint_as_float(v[j]), pr[j * 128]));
      }
      tcgen05_fence_before();
      mbar_arrive(&tempty[acc]);  // the accumulator and the partial sums are in registers
#pragma unroll
      for (int r = 1; r < KS; r++) mbar_arrive_remote(pempty, (uint32_t)r);
      EP::chunk32_with(rp, v, pre, m, n0, p.M, ebase, eaccs, nullptr, 0, -1, nullptr, nullptr, -1);
      // publish h_i: each of the 128 threads counts itself in once ITS stores are visible at device
      // scope (release) -- no fence + barrier + single store in series; step i is complete at 128 (i + 1)
      if (threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 6);
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(cp.flags + tn * p.num_m + tm), "r"(1u) : "memory");
      if (threadIdx.x == EPI_WARP0 * 32) chain_stamp(cp, tile, i, 7);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TN) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// operand with `mn` rows/cols along MN and `k` along K.
//   mn_major: memory is [k][mn] (mn contiguous, leading dim ld): dims (mn, k), box (32, TK), 32-byte-atom swizzle
//   k_major:  memory is [mn][k] (k contiguous):                   dims (k, mn), box (TK, tile_mn), 128B swizzle
static int make_map(CUtensorMap* map, const void* ptr, long long mn, long long k, long long ld, bool mn_major,
                    int tile_mn, int eb) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) YOTA_FAIL(YOTA_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * eb};
  const cuuint32_t tk = 128 / eb;
  cuuint32_t box[2], estr[2] = {1, 1};
  CUtensorMapSwizzle sw;
  if (mn_major) {
    dims[0] = (cuuint64_t)mn;
    dims[1] = (cuuint64_t)k;
    box[0] = tk;
    box[1] = tk;
    sw = eb == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  } else {
    dims[0] = (cuuint64_t)k;
    dims[1] = (cuuint64_t)mn;
    box[0] = tk;
    box[1] = (cuuint32_t)tile_mn;
    sw = CU_TENSOR_MAP_SWIZZLE_128B;
  }
  CUresult r = enc(map, eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) YOTA_FAIL(YOTA_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return YOTA_OK;
}

// k-blocks (of 32 K-elements) per accumulation chunk of the compensated mode (YOTA_X3_KCHUNK
// overrides).  Measured on B200, C3 at full size against the Float64 oracle, worst error / max|g|:
//   1-CTA kernel (chunks added to C by the L2, red.global.add.f32: every chunk is a pass over C):
//     8 k-blocks 1.17e-5 (just over the 1e-5 gate), 4 k-blocks 7.0e-6, 2 k-blocks 4.5e-6 -- default 4;
//   2-CTA kernel (chunks summed in registers, free): 4 k-blocks 7.0e-6 at 35.2 evals/s, 2 k-blocks
//     4.5e-6 at 34.7, 1 k-block 3.5e-6 at 34.9 -- default 1.
static int x3_kchunk(int dflt) {
  const char* e = getenv("YOTA_X3_KCHUNK");
  const int v = e && *e ? atoi(e) : dflt;
  return v < 1 ? 1 : v;
}

struct Maps {
  CUtensorMap a, b, al, bl;  // al / bl: the lo shadows of the compensated mode (copies of a / b otherwise)
};

template <int TN, bool A_MN, bool B_MN, int EB, bool X3 = false, int EPI = -1, int ZIN = 0, int KS = 1>
static int launch(const GemmArgs& g, const Maps& mp, int sm_count, cudaStream_t s, const RegionParams* rp = nullptr) {
  static bool attr = false;
  auto kern = gemm_tc_kernel<TN, A_MN, B_MN, EB, X3, EPI, ZIN, KS>;
  constexpr int SMEM = Smem<TN, X3>::TOTAL + (KS > 1 ? (KS - 1) * TN * 128 * 4 : 0);
  static_assert(SMEM <= 227 * 1024, "split-K partial sums must fit beside the ring");
  if (!attr) {
    YOTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  Params p;
  p.M = g.M;
  p.N = g.N;
  p.K = g.K;
  p.C = g.C;
  p.ldc = g.ldc;
  p.accumulate = g.accumulate;
  p.spread = getenv("YOTA_TMA_SPREAD") ? 1 : 0;
  p.kchunk = x3_kchunk(4);
  p.tsh = nullptr;
  p.tsh_ld = 0;
  p.tsh_out = -1;
  p.bsh = nullptr;
  p.bsh_out = -1;
  p.rs_out = nullptr;
  p.rs_scale = 1.f;
  p.rs_cnt = nullptr;
  p.early_a = g.pdl && g.early_a;
  p.ksplit = 1;
  p.part = nullptr;
  p.num_m = (int)(g.M / TM);
  p.num_n = (int)(g.N / TN);
  const int tiles = p.num_m * p.num_n;
  static const RegionParams no_rp{};
  const int clusters = tiles < sm_count / KS ? tiles : sm_count / KS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = KS == 1 ? dim3(tiles < sm_count ? tiles : sm_count) : dim3(KS * (clusters < 1 ? 1 : clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (KS > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = KS;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    na++;
  }
  if (g.pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  YOTA_CUDA(cudaLaunchKernelEx(&cfg, kern, mp.a, mp.b, mp.al, mp.bl, p, rp ? *rp : no_rp));
  return YOTA_OK;
}


// =============================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x 256 tile.  Each CTA
// holds its own 128 rows of A and HALF of B (128 columns); the tensor cores of both SMs read
// both halves of B, so per-CTA L2->smem traffic per MMA drops by a third (32 KiB instead of
// 48 KiB per 128-byte k-block) and the ring is 6 stages deep instead of 4.  At 128x256x32 the
// single-CTA kernel pulls ~17 TB/s through L2 -- that, not the tensor pipe, is its ceiling
// (profiles/README.md).  Protocol (leader = CTA rank 0 of the pair):
//   full[s]   lives in the leader; TMA loads of BOTH CTAs complete_tx on it (cta_group::2 form),
//             the peer's producer also arrives on it remotely (count 2)
//   empty[s]  per CTA; the leader's tcgen05.commit multicasts the arrive to both CTAs
//   tfull[a]  per CTA, multicast commit;  tempty[a] in the leader, 256 epilogue threads arrive
// =============================================================================================
constexpr int STAGES2 = 6;
constexpr int TN2 = 256;

template <bool X3>
struct Smem2T {
  static constexpr int NST = X3 ? 3 : STAGES2;
  static constexpr int A_BYTES = TM * 128;
  static constexpr int B_BYTES = (TN2 / 2) * 128;
  static constexpr int HALF = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = X3 ? 2 * HALF : HALF;
  static constexpr int BAR_OFF = NST * STAGE_BYTES;
  static constexpr int TSH_OFF = BAR_OFF + 256;           // 4 epilogue warps x (32 x 32 floats): transpose staging
  static constexpr int TOTAL = TSH_OFF + 4 * 4096 + 1024;  // + slack for 1 KiB alignment
};

// TMA load whose completion is signalled on the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// arrive (when the MMAs issued so far have completed) on the barrier at this smem offset in every
// CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// TMA load delivered to the same smem offset of every CTA in `cta_mask`; each copy's completion is
// signalled on the mbarrier (same offset) of the leader of the receiving CTA's pair
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                   uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
template <int EB>
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  if (EB == 4)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}



// MC = 2: clusters of FOUR CTAs -- two pairs computing the tiles (m, 2j) and (m, 2j + 1), which
// read the same 256 rows of A.  Each CTA fetches only half of its 128-row A box and TMA-multicasts
// it to the CTA holding the same rows in the other pair: per k-block a CTA pulls 8 + 16 KiB through
// L2 instead of 16 + 16.  At 256 x 256 x 32 tiles with Float32 operands the L2 -> SM operand stream
// (64 KiB per 4.2 MFLOP) is what bounds the kernel once the batch per GPU is small (N = 1024:
// 520 TFLOP/s at 7.8 TB/s of L2 reads).  An smem stage is then written by TMA loads of BOTH
// pairs, so empty[s] collects the commits of both pairs' MMA issuers (count MC).
// Compensated mode: EIGHT epilogue warps (two per TMEM lane quarter, 128 columns each), because the
// chunks of a tile are summed in REGISTERS -- 128 running sums per thread -- instead of being added
// to C by the L2 (red.global.add.f32): at one chunk per 128 K-elements that was K / 128 read-modify-
// write passes over C per GEMM (8.6 GB of L2 atomics for 8192 x 8192 x 4096: the kernel was bound by
// them, 1.7 ms where three TF32 passes take 1.3).
constexpr int THREADS_X3 = 320;
template <bool A_MN, bool B_MN, int EB, int EPI, int ZIN, bool X3, int MC>
__global__ void __launch_bounds__(X3 ? THREADS_X3 : THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_al, const __grid_constant__ CUtensorMap map_bl, const Params p,
                const __grid_constant__ RegionParams rp) {
  static_assert(!X3 || EB == 4, "the compensated mode splits Float32 operands");
  using S = Smem2T<X3>;
  constexpr int STAGES2 = S::NST;  // shadows the namespace constant: 3 deep with the lo tiles aboard
  using E = El<EB>;
  constexpr int TK = E::TK;
  constexpr int TNH = TN2 / 2;  // B columns held by one CTA
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES2;
  uint64_t* tfull = empty + STAGES2;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;               // position inside the CTA pair
  // MC = 2 launches ask for clusters of 4 as the PREFERRED size over regular clusters of 2
  // (cudaLaunchAttributePreferredClusterDimension): the GPCs of a B200 hold an odd number of TPCs,
  // so only 33 four-SM clusters fit on the 148 SMs -- the 8 TPCs left over run as regular pairs,
  // which fetch all of their A rows themselves.  A group of 4 consecutive CTAs is either one
  // preferred cluster or two regular ones and does the same work either way (static schedule).
  uint32_t csize;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int mc = MC > 1 ? (int)(csize >> 1) : 1;            // pairs in THIS cluster
  const uint32_t pp = MC > 1 ? (blockIdx.x & 3u) >> 1 : 0u;  // pair index inside the group: which of the MC tiles
  const uint32_t cpp = crank >> 1;                          // pair index inside the cluster
  const uint32_t lead = crank & ~1u;              // cluster rank of this pair's leader
  const bool leader = rank == 0;
  const int pair = blockIdx.x / (2 * MC), num_pairs = gridDim.x / (2 * MC);  // work units: MC tiles adjacent along N
  const int ksplit = (EPI < 0 && !X3) ? p.ksplit : 1;
  const int base_tiles = p.num_m * (p.num_n / MC);
  const int num_tiles = base_tiles * ksplit;  // K slices are scheduled like extra tiles
  const int nkb = (int)(p.K / TK) / ksplit;
  const uint16_t mask_pair = (uint16_t)(3u << (2 * cpp)), mask_all = (uint16_t)((1u << (2 * mc)) - 1u);
  const uint16_t mask_same_rows = (uint16_t)((1u << rank) | (1u << (2 + rank)));  // MC = 2: this CTA and its twin

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; s++) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], mc);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], X3 ? 2 * (THREADS_X3 - 64) : 256);  // the epilogue threads of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * TN2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer warp (both CTAs) =====================
    constexpr int NA = A_MN ? TM / E::CHUNK : 1;
    constexpr int NB = B_MN ? TNH / E::CHUNK : 1;
    int stage = 0;
    uint32_t phase = 0;
    constexpr int NAI = NA / MC > 0 ? NA / MC : 1;  // A boxes this CTA issues itself (MN-major A: half the chunks)
    for (int tk = pair; tk < num_tiles; tk += num_pairs) {
      const int ks = tk / base_tiles, t = tk - ks * base_tiles;
      const int m0 = (t % p.num_m) * (2 * TM) + (int)rank * TM;
      const int n0 = ((t / p.num_m) * MC + (int)pp) * TN2 + (int)rank * TNH;
      for (int kb = 0; kb < nkb; kb++) {
        if (lane == 0) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader)
            mbar_expect_tx(&full[stage], 2 * S::STAGE_BYTES);
          else
            mbar_arrive_remote(&full[stage], lead);
        }
        __syncwarp();
        const int k0 = (ks * nkb + kb) * TK;
#pragma unroll
        for (int part = 0; part < (X3 ? 2 : 1); part++) {  // part 1: the lo tiles of the compensated mode
          unsigned char* sa = smem + stage * S::STAGE_BYTES + part * S::HALF;
          unsigned char* sb = sa + S::A_BYTES;
          const CUtensorMap* ma = part ? &map_al : &map_a;
          const CUtensorMap* mb = part ? &map_bl : &map_b;
          const int l0 = part * (NAI + NB);
          // A box c of this CTA (MC = 2: it issues boxes [pp * NAI, (pp + 1) * NAI) of the NA -- or
          // rows [pp * 64, +64) of a K-major box -- and multicasts them to its twin in the other pair)
          auto load_a = [&](int i) {
            if constexpr (MC == 1) {
              if (A_MN)
                tma_load_2d_2sm(sa + i * E::CHUNK_BYTES, ma, &full[stage], m0 + E::CHUNK * i, k0);
              else
                tma_load_2d_2sm(sa, ma, &full[stage], k0, m0);
            } else if (mc == 2) {  // preferred cluster: half of the boxes, multicast to the twin CTA
              if (A_MN) {
                const int c = (int)pp * NAI + i;
                tma_load_2d_2sm_mc(sa + c * E::CHUNK_BYTES, ma, &full[stage], m0 + E::CHUNK * c, k0, mask_same_rows);
              } else
                tma_load_2d_2sm_mc(sa + pp * (S::A_BYTES / MC), ma, &full[stage], k0, m0 + (int)pp * (TM / MC),
                                   mask_same_rows);
            } else {  // regular pair of a multicast-capable launch: both halves, no multicast
#pragma unroll
              for (int h = 0; h < MC; h++) {
                if (A_MN) {
                  const int c = h * NAI + i;
                  tma_load_2d_2sm(sa + c * E::CHUNK_BYTES, ma, &full[stage], m0 + E::CHUNK * c, k0);
                } else
                  tma_load_2d_2sm(sa + h * (S::A_BYTES / MC), ma, &full[stage], k0, m0 + h * (TM / MC));
              }
            }
          };
          auto load_b = [&](int c) {
            if (B_MN)
              tma_load_2d_2sm(sb + c * E::CHUNK_BYTES, mb, &full[stage], n0 + E::CHUNK * c, k0);
            else
              tma_load_2d_2sm(sb, mb, &full[stage], k0, n0);
          };
          if (!p.spread) {
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < NAI; i++) load_a(i);
#pragma unroll
              for (int c = 0; c < NB; c++) load_b(c);
            }
          } else if (lane >= l0 && lane < l0 + NAI) {
            load_a(lane - l0);
          } else if (lane >= l0 + NAI && lane < l0 + NAI + NB) {
            load_b(lane - l0 - NAI);
          }
        }
        if (++stage == STAGES2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA drives both tensor cores =====================
    if (leader && lane == 0) {
      constexpr uint32_t FMT = EB == 4 ? 2u : 1u;
      const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TN2 >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int kc = X3 ? p.kchunk : nkb;
      for (int t = pair; t < num_tiles; t += num_pairs)
       for (int ch = 0; ch < (X3 ? (nkb + kc - 1) / kc : 1); ch++) {  // one accumulator per chunk of k-blocks (X3) / per tile
        const int kb0 = X3 ? ch * kc : 0;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN2;
        const int kend = kb0 + kc < nkb ? kb0 + kc : nkb;
        for (int kb = kb0; kb < kend; kb++) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t da = A_MN ? make_desc(sa + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                     : make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = B_MN ? make_desc(sb + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                     : make_desc(sb + k * 32, 16, 1024, 2);
            umma_2sm<EB>(d_tmem, da, db, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            if constexpr (X3) {
              const uint32_t sal = sa + S::HALF, sbl = sal + S::A_BYTES;
              const uint64_t dal = A_MN ? make_desc(sal + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                        : make_desc(sal + k * 32, 16, 1024, 2);
              const uint64_t dbl = B_MN ? make_desc(sbl + k * E::MN_KSTEP, E::CHUNK_BYTES, E::MN_SBO, E::MN_LAYOUT)
                                        : make_desc(sbl + k * 32, 16, 1024, 2);
              umma_2sm<EB>(d_tmem, da, dbl, idesc, 1u);
              umma_2sm<EB>(d_tmem, dal, db, idesc, 1u);
            }
          }
          umma_commit_2sm(&empty[stage], mask_all);  // every CTA whose TMA loads write into these smem slots
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull[acc], mask_pair);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if constexpr (X3) {
    // ===================== epilogue warps, compensated mode: chunk sums in registers =====================
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int hc = (warp - EPI_WARP0) >> 2;  // which 128 of the tile's 256 columns
    int acc = 0;
    uint32_t acc_phase = 0;
    const int kc = p.kchunk, nch = (nkb + kc - 1) / kc;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const long long m = (long long)(t % p.num_m) * (2 * TM) + (long long)rank * TM + q * 32 + lane;
      const long long tile_n = (long long)(t / p.num_m) * MC + pp;
      const long long n0 = tile_n * TN2 + hc * (TN2 / 2);
      using EP = Epilogue<(EPI >= 0 ? EPI : 0), ZIN>;
      float ebase[EP::NS], eaccs[EP::NO];
      if constexpr (EPI >= 0) EP::begin_tile(rp, m, n0, p.M, ebase, eaccs, num_tiles <= num_pairs);
      float sums[TN2 / 2];
#pragma unroll
      for (int j = 0; j < TN2 / 2; j++) sums[j] = 0.f;
      for (int ch = 0; ch < nch; ch++) {
        mbar_wait(&tfull[acc], acc_phase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < TN2 / 32; c++) {
          uint32_t v[16];
          const uint32_t taddr = tmem_base + (uint32_t)(acc * TN2 + hc * (TN2 / 2) + c * 16) + ((uint32_t)(q * 32) << 16);
          tmem_ld16(taddr, v);
#pragma unroll
          for (int j = 0; j < 16; j++) sums[c * 16 + j] = __fadd_rn(sums[c * 16 + j], __uint_as_float(v[j]));  // round to nearest
        }
        tcgen05_fence_before();
        mbar_arrive_remote(&tempty[acc], lead);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      float* crow = p.C + m + n0 * p.ldc;
#pragma unroll
      for (int c = 0; c < TN2 / 64; c++) {
        if constexpr (EPI >= 0) {
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; j++) z[j] = __float_as_uint(sums[c * 32 + j]);
          EP::chunk32(rp, z, m, n0 + c * 32, p.M, ebase, eaccs, nullptr, 0, -1, nullptr, nullptr, -1);
        } else {
          float* cp = crow + (long long)(c * 32) * p.ldc;
          if (p.accumulate) {
#pragma unroll
            for (int h = 0; h < 4; h++) {  // 8 old values in flight at a time: the 128 sums leave few registers
              float old[8];
#pragma unroll
              for (int j = 0; j < 8; j++) old[j] = __ldcg(cp + (h * 8 + j) * p.ldc);
#pragma unroll
              for (int j = 0; j < 8; j++) cp[(h * 8 + j) * p.ldc] = __fadd_rn(sums[c * 32 + h * 8 + j], old[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) cp[j * p.ldc] = sums[c * 32 + j];
          }
        }
      }
      if constexpr (EPI >= 0) {
        // row-sum sinks: the thread holding columns 128..255 of this row hands its partial sums to
        // the one holding 0..127 (fixed order: left half + right half), which delivers them
        float* xch = reinterpret_cast<float*>(smem + S::TSH_OFF);  // [NO][128 rows]
        if (hc == 1) {
#pragma unroll
          for (int k = 0; k < EP::NO; k++) xch[k * 128 + q * 32 + lane] = eaccs[k];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (hc == 0) {
#pragma unroll
          for (int k = 0; k < EP::NO; k++) eaccs[k] = __fadd_rn(eaccs[k], xch[k * 128 + q * 32 + lane]);
          EP::end_tile(rp, m, tile_n, p.M, eaccs, p, reinterpret_cast<unsigned*>(smem + S::BAR_OFF + 200));
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");  // xch is reused by the next tile
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs): own 128 rows x 256 columns =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tk = pair; tk < num_tiles; tk += num_pairs) {
      const int ks = tk / base_tiles, t = tk - ks * base_tiles;
      const long long m = (long long)(t % p.num_m) * (2 * TM) + (long long)rank * TM + q * 32 + lane;
      const long long tile_n = (long long)(t / p.num_m) * MC + pp;
      const long long n0 = tile_n * TN2;
      const long long ldc = ksplit > 1 ? p.M : p.ldc;
      float* crow = (ksplit > 1 ? p.part + (long long)ks * p.M * p.N : p.C) + m + n0 * ldc;
      using EP = Epilogue<(EPI >= 0 ? EPI : 0), ZIN>;
      float ebase[EP::NS], eaccs[EP::NO];
      if constexpr (EPI >= 0) EP::begin_tile(rp, m, n0, p.M, ebase, eaccs, num_tiles <= num_pairs);
      const bool add_old = EPI < 0 && p.accumulate && ksplit == 1;
      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < TN2 / 32; c++) {
        uint32_t v[32];
        float old[32];
        if (add_old) {  // see gemm_tc_kernel: old values of C in flight before the accumulator read
          const float* cp0 = crow + (long long)(c * 32) * ldc;
#pragma unroll
          for (int j = 0; j < 32; j++) old[j] = __ldcg(cp0 + j * ldc);
        }
        const uint32_t taddr = tmem_base + (uint32_t)(acc * TN2 + c * 32) + ((uint32_t)(q * 32) << 16);
        tmem_ld32(taddr, v);
        if constexpr (EPI >= 0) {
          EP::chunk32(rp, v, m, n0 + c * 32, p.M, ebase, eaccs, p.tsh, p.tsh_ld, p.tsh_out,
                      reinterpret_cast<float4*>(smem + S::TSH_OFF) + q * 256, p.bsh, p.bsh_out);
        } else {
          float* cp = crow + (long long)(c * 32) * ldc;
          if (add_old) {
#pragma unroll
            for (int j = 0; j < 32; j++) cp[j * ldc] = __uint_as_float(v[j]) + old[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) cp[j * ldc] = __uint_as_float(v[j]);
          }
        }
      }
      if constexpr (EPI >= 0)
        EP::end_tile(rp, m, tile_n, p.M, eaccs, p, reinterpret_cast<unsigned*>(smem + S::BAR_OFF + 200));
      tcgen05_fence_before();
      mbar_arrive_remote(&tempty[acc], lead);  // the leader's MMA warp waits for both CTAs' epilogues
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's smem / TMEM may still be read by the leader's MMAs until here
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TN2) : "memory");
  }
}

template <bool A_MN, bool B_MN, int EB, int EPI, int ZIN, bool X3, int MC>
static int launch2_mc(const GemmArgs& g, const Maps& mp, int sm_count, cudaStream_t s, const RegionParams* rp) {
  static bool attr = false;
  static int max_clusters = 0;  // co-resident clusters of 2 * MC CTAs the device can hold
  auto kern = gemm_tc2_kernel<A_MN, B_MN, EB, EPI, ZIN, X3, MC>;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(X3 ? THREADS_X3 : THREADS);
  cfg.dynamicSmemBytes = Smem2T<X3>::TOTAL;
  cfg.stream = s;
  // regular clusters are CTA pairs; MC = 2 prefers clusters of four (two pairs sharing A by
  // multicast) wherever the GPC has room for them
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributePreferredClusterDimension;
  at[1].val.preferredClusterDim.x = 2 * MC;
  at[1].val.preferredClusterDim.y = 1;
  at[1].val.preferredClusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = MC > 1 && !getenv("YOTA_GEMM_STRICT_CLUSTER4") ? 2 : 1;
  if (cfg.numAttrs == 1) at[0].val.clusterDim.x = 2 * MC;
  if (!attr) {
    YOTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2T<X3>::TOTAL));
    cfg.gridDim = dim3(2 * MC);
    int n = 0;
    const int numAttrs = cfg.numAttrs;
    cfg.numAttrs = 1;  // co-resident REGULAR clusters (pairs, or fours with YOTA_GEMM_STRICT_CLUSTER4)
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0)
      max_clusters = numAttrs == 2 ? n / MC : n;  // groups of MC pairs
    cfg.numAttrs = numAttrs;
    (void)cudaGetLastError();
    attr = true;
  }
  Params p;
  p.M = g.M;
  p.N = g.N;
  p.K = g.K;
  p.C = g.C;
  p.ldc = g.ldc;
  p.accumulate = g.accumulate;
  p.spread = getenv("YOTA_TMA_SPREAD") ? 1 : 0;
  p.kchunk = x3_kchunk(1);
  p.tsh = g.tsh;
  p.tsh_ld = g.tsh_ld;
  p.tsh_out = g.tsh_out;
  p.bsh = static_cast<__nv_bfloat16*>(g.bsh);
  p.bsh_out = g.bsh_out;
  p.rs_out = g.rs_out;
  p.rs_scale = g.rs_scale;
  p.rs_cnt = g.rs_cnt;
  p.early_a = 0;
  p.ksplit = (EPI < 0 && !X3 && g.ksplit > 1) ? g.ksplit : 1;
  p.part = static_cast<float*>(g.splitk_ws);
  p.num_m = (int)(g.M / (2 * TM));
  p.num_n = (int)(g.N / TN2);
  const int units = p.num_m * (p.num_n / MC) * p.ksplit;  // one unit = MC tiles adjacent along N = one cluster's work item
  int clusters = sm_count / (2 * MC);
  if (max_clusters > 0 && clusters > max_clusters) clusters = max_clusters;
  if (units < clusters) clusters = units;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3(2 * MC * clusters);
  static const RegionParams no_rp{};
  YOTA_CUDA(cudaLaunchKernelEx(&cfg, kern, mp.a, mp.b, mp.al, mp.bl, p, rp ? *rp : no_rp));
  return YOTA_OK;
}

// tensor maps of both operands (mode 3: also of their lo shadows, same dims and strides);
// a_rows: rows of A per TMA box of a K-major A tile (TM, or TM / 2 when two pairs share A)
static int make_maps(Maps* mp, const GemmArgs& g, bool A_MN, bool B_MN, int tile_n, int a_rows = TM) {
  const int eb = g.mode == 2 ? 2 : 4;
  YOTA_TRY(make_map(&mp->a, g.A, g.M, g.K, g.lda, A_MN, a_rows, eb));
  YOTA_TRY(make_map(&mp->b, g.B, g.N, g.K, g.ldb, B_MN, tile_n, eb));
  if (g.mode == 3) {
    if (!g.A_lo || !g.B_lo) YOTA_FAIL(YOTA_E_ARG, "internal: compensated GEMM without lo shadows");
    YOTA_TRY(make_map(&mp->al, g.A_lo, g.M, g.K, g.lda, A_MN, a_rows, eb));
    YOTA_TRY(make_map(&mp->bl, g.B_lo, g.N, g.K, g.ldb, B_MN, tile_n, eb));
  } else {
    mp->al = mp->a;
    mp->bl = mp->b;
  }
  return YOTA_OK;
}

// two pairs per cluster share their A rows whenever the tiles pair up along N
static bool use_mc2(const GemmArgs& g) { return (g.N / TN2) % 2 == 0 && !getenv("YOTA_GEMM_NO_MULTICAST"); }

template <bool A_MN, bool B_MN, int EB, int EPI = -1, int ZIN = 0, bool X3 = false>
static int launch2(const GemmArgs& g, int sm_count, cudaStream_t s, const RegionParams* rp = nullptr) {
  Maps mp;
  if (use_mc2(g)) {
    YOTA_TRY(make_maps(&mp, g, A_MN, B_MN, TN2 / 2, TM / 2));
    return launch2_mc<A_MN, B_MN, EB, EPI, ZIN, X3, 2>(g, mp, sm_count, s, rp);
  }
  YOTA_TRY(make_maps(&mp, g, A_MN, B_MN, TN2 / 2, TM));
  return launch2_mc<A_MN, B_MN, EB, EPI, ZIN, X3, 1>(g, mp, sm_count, s, rp);
}

}  // namespace tc

// SMs a persistent GEMM grid may occupy right now (see yota_ctx::sm_reserved)
static inline int grid_sms(const yota_ctx* ctx) {
  const char* e = getenv("YOTA_GEMM_GRID_SMS");  // diagnostic override (tests/occupy_probe.cu)
  const int n = e && *e ? atoi(e) : ctx->sm_count - ctx->sm_reserved;
  return n < 2 ? 2 : n;
}

bool gemm_tc_eligible(const GemmArgs& g) {
  if (g.mode < 1 || g.mode > 3) return false;
  const int tk = g.mode == 2 ? 64 : 32;
  if (g.M < tc::TM || g.N < 128 || g.K < tk) return false;
  if (g.M % tc::TM || g.N % 128 || g.K % tk) return false;
  if ((g.lda | g.ldb) % 8) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B)) & 15) return false;
  return true;
}

bool gemm_tc2_eligible(const GemmArgs& g, int sm_count) {
  if (!gemm_tc_eligible(g) || getenv("YOTA_GEMM_1CTA")) return false;
  if (g.M % 256 || g.N % 256) return false;
  const long long tiles = (g.M / 256) * (g.N / 256);
  return getenv("YOTA_GEMM_FORCE_2CTA") ? tiles >= 1 : tiles >= sm_count / 4;
}

// Few 256 x 256 output tiles and a long K: cut K so that about 64 CTA pairs have work.  The
// slices are summed in slice order afterwards -- deterministic, one more 4-byte pass over
// ksplit + 1 copies of C, nothing next to the operand streams of such a contraction.
int gemm_tc2_ksplit(const GemmArgs& g, int sm_count) {
  if (g.mode != 1 || getenv("YOTA_GEMM_1CTA") || getenv("YOTA_GEMM_NO_KSPLIT")) return 0;
  if (!gemm_tc_eligible(g) || g.M % 256 || g.N % 256 || g.ldc != g.M) return 0;
  const long long tiles = (g.M / 256) * (g.N / 256);
  if (tiles >= sm_count / 4) return 0;  // enough tiles for the plain 2-CTA schedule
  int ks = 0;
  for (int c = 2; c <= 8; c *= 2)
    if (tiles * c <= sm_count / 2 && (g.K / 32) % c == 0 && g.K / c >= 2048) ks = c;
  return ks;
}

__global__ void splitk_reduce_kernel(float4* __restrict__ C, const float4* __restrict__ part, long long n4, int ks,
                                     int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = accumulate ? C[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < ks; k++) {  // slice order: the same sum on every run
      const float4 v = __ldcs(part + (long long)k * n4 + i);
      a.x = __fadd_rn(a.x, v.x);
      a.y = __fadd_rn(a.y, v.y);
      a.z = __fadd_rn(a.z, v.z);
      a.w = __fadd_rn(a.w, v.w);
    }
    C[i] = a;
  }
}

// the accumulator may stand in for input 0 or 1 of the chain, if that input is a FULL operand
template <int ID>
static constexpr bool epi_z_ok(int z_in) {
  return z_in >= 0 && z_in < 2 && z_in < ChainOf<ID>::value.n_in && ChainOf<ID>::value.in_cls[z_in] == CLS_FULL;
}

bool gemm_epilogue_chain_ok(int static_id, int transA, int transB, int z_in) {
  // instantiated: forward (A not transposed, B not transposed) and dh (A transposed) GEMMs
  if (transB) return false;
  switch (static_id) {
#define X(ID) \
  case ID:    \
    return epi_z_ok<ID>(z_in);
    YOTA_EPI_CHAIN_LIST(X)
#undef X
  }
  return false;
}

namespace tc {
template <int ID, int ZIN>
static int launch_epi(const GemmArgs& g, int sm_count, cudaStream_t s, const RegionParams& rp, bool A_MN) {
  if constexpr (epi_z_ok<ID>(ZIN)) {
#define GOE(a)                                                                           \
  return g.mode == 3   ? launch2<a, false, 4, ID, ZIN, true>(g, sm_count, s, &rp)    \
         : g.mode == 1 ? launch2<a, false, 4, ID, ZIN>(g, sm_count, s, &rp)          \
                       : launch2<a, false, 2, ID, ZIN>(g, sm_count, s, &rp)
    if (A_MN) GOE(true);
    GOE(false);
#undef GOE
  } else {
    YOTA_FAIL(YOTA_E_ARG, "internal: chain %d has no FULL operand %d", ID, ZIN);
  }
}
}  // namespace tc

// tile width the 1-CTA kernel uses: the widest that still gives half the SMs a CTA; skinny outputs
// go down to 128 x 32 (64 for BF16 operands, whose MN-major boxes are 64 elements wide)
static int pick_tn(const GemmArgs& g, int sm_count) {
  const int eb = g.mode == 2 ? 2 : 4;
  const int widths[4] = {256, 128, 64, 32};
  for (int w : widths)
    if (w >= (eb == 4 ? 32 : 64) && g.N % w == 0 && (g.M / tc::TM) * (g.N / w) >= sm_count / 2) return w;
  return eb == 4 ? 32 : 64;
}

// store-only chains that also exist as epilogues of the 1-CTA kernel (TF32, 128 x 32 tiles)
template <int ID>
static constexpr bool epi1_chain_stores_only() {
  for (int k = 0; k < ChainOf<ID>::value.n_out; k++)
    if (ChainOf<ID>::value.out_kind[k] != OUT_STORE) return false;
  return true;
}
bool gemm_epilogue1_chain_ok(const GemmArgs& g, int sm_count, int static_id, int z_in) {
  if (g.mode != 1 || !gemm_tc_eligible(g) || gemm_tc2_eligible(g, sm_count) || g.transB || g.accumulate) return false;
  if (pick_tn(g, sm_count) != 32) return false;
  switch (static_id) {
#define X(ID) \
  case ID:    \
    return epi_z_ok<ID>(z_in) && epi1_chain_stores_only<ID>();
    YOTA_EPI1_CHAIN_LIST(X)
#undef X
  }
  return false;
}

// skinny output with a long K: a cluster of 4 CTAs per tile, each streaming a quarter of K
static bool use_splitk4(const GemmArgs& g, int sm_count) {
  if (g.mode != 1 || g.transB || getenv("YOTA_GEMM_NO_SPLITK")) return false;
  const long long tiles = (g.M / tc::TM) * (g.N / 32), nkb = g.K / 32;
  return pick_tn(g, sm_count) == 32 && tiles <= sm_count / 4 && nkb % 4 == 0 && nkb >= 8;
}

namespace tc {
template <int ID, int ZIN>
static int launch_epi1(const GemmArgs& g, const Maps& mp, int sm_count, cudaStream_t s, const RegionParams& rp, bool A_MN) {
  if constexpr (epi_z_ok<ID>(ZIN) && epi1_chain_stores_only<ID>()) {
    if (use_splitk4(g, sm_count)) {
      if (A_MN) return launch<32, true, false, 4, false, ID, ZIN, 4>(g, mp, sm_count, s, &rp);
      return launch<32, false, false, 4, false, ID, ZIN, 4>(g, mp, sm_count, s, &rp);
    }
    if (A_MN) return launch<32, true, false, 4, false, ID, ZIN>(g, mp, sm_count, s, &rp);
    return launch<32, false, false, 4, false, ID, ZIN>(g, mp, sm_count, s, &rp);
  } else {
    YOTA_FAIL(YOTA_E_ARG, "internal: chain %d cannot run as a 1-CTA epilogue", ID);
  }
}
}  // namespace tc

namespace tc {
template <bool A_MN, int ID, int ZIN>
static int launch_chain_t(const GemmArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, const ChainParams& cp,
                          const RegionParams& rp, cudaStream_t s) {
  if constexpr (epi_z_ok<ID>(ZIN) && epi1_chain_stores_only<ID>()) {
    static bool attr = false;
    static int max_clusters = 0;
    auto kern = gemm_chain_kernel<A_MN, ID, ZIN>;
    constexpr int SMEM = Smem<32, false>::TOTAL + 3 * 32 * 128 * 4;
    const int tiles = (int)((g.M / TM) * (g.N / 32));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * tiles);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (!attr) {
      YOTA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess) max_clusters = n;
      (void)cudaGetLastError();
      attr = true;
    }
    if (tiles > max_clusters) return 1;  // the clusters spin on each other's flags: all must be resident
    Params p{};
    p.M = g.M;
    p.N = g.N;
    p.K = g.K;
    p.C = g.C;
    p.ldc = g.ldc;
    p.kchunk = 1;
    p.tsh_out = -1;
    p.bsh_out = -1;
    p.rs_scale = 1.f;
    p.ksplit = 1;
    p.num_m = (int)(g.M / TM);
    p.num_n = (int)(g.N / 32);
    YOTA_CUDA(cudaMemsetAsync(cp.flags, 0, (size_t)tiles * sizeof(unsigned), s));
    YOTA_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p, cp, rp));
    return YOTA_OK;
  } else {
    return 1;
  }
}
}  // namespace tc

// YOTA_CHAIN_TRACE=1: device buffer for the time stamps of the last persistent recurrence launched
static unsigned long long* g_chain_trace = nullptr;
static int g_chain_trace_steps = 0;
static unsigned long long* chain_trace_buffer(int nsteps) {
  if (!getenv("YOTA_CHAIN_TRACE") || nsteps > 4096) return nullptr;
  if (!g_chain_trace && cudaMalloc(&g_chain_trace, 4096 * 8 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
  g_chain_trace_steps = nsteps;
  return g_chain_trace;
}
// diagnostic export (not part of include/yota_cuda.h): copies the stamps of the last traced recurrence
extern "C" int yota_debug_chain_trace(unsigned long long* out, int cap_steps) {
  if (!g_chain_trace) return 0;
  const int n = g_chain_trace_steps < cap_steps ? g_chain_trace_steps : cap_steps;
  if (cudaMemcpy(out, g_chain_trace, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

// The whole recurrence  out_i = chain(A * B_i, operands_i),  B_{i+1} = out_i,  i < nsteps, as one
// persistent launch (tc::gemm_chain_kernel).  g0 / rp0: arguments of step 0; b_delta: floats from
// B_i to B_{i+1}; e_dcol: columns from the chain's FULL operands / stores of step i to those of step
// i + 1; flags: >= tiles zero-initialisable unsigneds owned by this chain.  Returns YOTA_OK when the
// recurrence was launched, 1 when it cannot run this way (the caller launches the steps one by
// one), a negative status on errors.
int launch_gemm_chain(yota_ctx* ctx, const GemmArgs& g0, int static_id, const RegionParams& rp0, int z_in, int nsteps,
                      long long b_delta, long long e_dcol, unsigned* flags, cudaStream_t s) {
  if (getenv("YOTA_NO_CHAIN") || nsteps < 2 || g0.mode != 1 || g0.accumulate || g0.transB || g0.M != g0.K) return 1;
  if (!gemm_epilogue1_chain_ok(g0, ctx->sm_count, static_id, z_in) || !use_splitk4(g0, ctx->sm_count)) return 1;
  const long long nkb = g0.K / 32 / 4;
  if (nkb > tc::Smem<32, false>::NSTAGE) return 1;  // the CTA's K-slice of A stays in the ring
  if (b_delta % g0.ldb) return 1;
  const long long b_dcol = b_delta / g0.ldb, span = (nsteps - 1) * (b_dcol < 0 ? -b_dcol : b_dcol);
  if (g0.N + span > 0x7fffffff || e_dcol * (nsteps - 1) > 0x7fffffff || e_dcol * (nsteps - 1) < -0x7fffffff) return 1;
  const float* b_base = static_cast<const float*>(g0.B) + (b_delta < 0 ? (nsteps - 1) * b_delta : 0);
  if (reinterpret_cast<uintptr_t>(b_base) & 15) return 1;
  const bool A_MN = !g0.transA;
  CUtensorMap ma, mb;
  YOTA_TRY(tc::make_map(&ma, g0.A, g0.M, g0.K, g0.lda, A_MN, tc::TM, 4));
  YOTA_TRY(tc::make_map(&mb, b_base, g0.N + span, g0.K, g0.ldb, false, 32, 4));
  tc::ChainParams cp{};
  cp.nsteps = nsteps;
  cp.b_col0 = (int)(b_delta < 0 ? span : 0);
  cp.b_dcol = (int)b_dcol;
  cp.e_dcol = e_dcol;
  cp.flags = flags;
  cp.err = ctx->err_dev;
  cp.trace = chain_trace_buffer(nsteps);
  switch (static_id) {
#define X(ID)                                                                                           \
  case ID:                                                                                              \
    if (z_in == 0)                                                                                      \
      return A_MN ? tc::launch_chain_t<true, ID, 0>(g0, ma, mb, cp, rp0, s) : tc::launch_chain_t<false, ID, 0>(g0, ma, mb, cp, rp0, s); \
    return A_MN ? tc::launch_chain_t<true, ID, 1>(g0, ma, mb, cp, rp0, s) : tc::launch_chain_t<false, ID, 1>(g0, ma, mb, cp, rp0, s);
    YOTA_EPI1_CHAIN_LIST(X)
#undef X
  }
  return 1;
}

// GEMM whose output tile goes straight through the consumer region's micro-program
int launch_gemm_tc_epilogue(yota_ctx* ctx, const GemmArgs& g, int static_id, const RegionParams& rp, int z_in,
                            cudaStream_t s) {
  const bool A_MN = !g.transA;
  if (gemm_epilogue1_chain_ok(g, ctx->sm_count, static_id, z_in)) {  // skinny output: 1-CTA kernel, 128 x 32 tiles
    tc::Maps mp;
    YOTA_TRY(tc::make_maps(&mp, g, A_MN, false, 32));
    switch (static_id) {
#define X(ID)                                                                        \
  case ID:                                                                           \
    return z_in == 0 ? tc::launch_epi1<ID, 0>(g, mp, grid_sms(ctx), s, rp, A_MN)    \
                     : tc::launch_epi1<ID, 1>(g, mp, grid_sms(ctx), s, rp, A_MN);
      YOTA_EPI1_CHAIN_LIST(X)
#undef X
    }
  }
  if (!gemm_tc2_eligible(g, ctx->sm_count) || g.accumulate ||
      !gemm_epilogue_chain_ok(static_id, g.transA, g.transB, z_in))
    YOTA_FAIL(YOTA_E_ARG, "internal: GEMM epilogue fusion planned for an ineligible GEMM");
  switch (static_id) {
#define X(ID)                                                                       \
  case ID:                                                                          \
    return z_in == 0 ? tc::launch_epi<ID, 0>(g, grid_sms(ctx), s, rp, A_MN)    \
                     : tc::launch_epi<ID, 1>(g, grid_sms(ctx), s, rp, A_MN);
    YOTA_EPI_CHAIN_LIST(X)
#undef X
  }
  YOTA_FAIL(YOTA_E_ARG, "internal: no epilogue kernel for chain %d", static_id);
}

// mode 1: A/B are the Float32 arrays (read as TF32).  mode 2: A/B point at BF16 shadows with the
// same dims / leading dimensions (see launch_cast_bf16 and the planner's shadow tensors).
// mode 3: A/B as in mode 1 plus A_lo/B_lo, the Float32 shadows x - tf32(x) (launch_split_tf32).
int launch_gemm_tc(yota_ctx* ctx, const GemmArgs& g, cudaStream_t s) {
  if (!gemm_tc_eligible(g)) YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "gemm: shape not eligible for the tensor-core kernel");
  const bool A_MN = !g.transA;      // A stored M x K column-major: M contiguous
  const bool B_MN = g.transB != 0;  // B stored N x K column-major: N contiguous
  const int eb = g.mode == 2 ? 2 : 4;
  tc::Maps mp;
  if (g.splitk_ws && g.ksplit > 1 && g.ksplit == gemm_tc2_ksplit(g, ctx->sm_count)) {
    int st;
    if (A_MN && B_MN) st = tc::launch2<true, true, 4>(g, grid_sms(ctx), s);
    else if (A_MN) st = tc::launch2<true, false, 4>(g, grid_sms(ctx), s);
    else if (B_MN) st = tc::launch2<false, true, 4>(g, grid_sms(ctx), s);
    else st = tc::launch2<false, false, 4>(g, grid_sms(ctx), s);
    YOTA_TRY(st);
    const long long n4 = g.M * g.N / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float4*>(g.C),
                                                          static_cast<const float4*>(g.splitk_ws), n4, g.ksplit,
                                                          g.accumulate);
    YOTA_CUDA(cudaGetLastError());
    return YOTA_OK;
  }
  const bool two_cta = gemm_tc2_eligible(g, ctx->sm_count);
  if (two_cta) {
    // each CTA of a pair loads a 128-row box of A and a 128-column box of B
#define GO2(a, b)                                                                     \
  return g.mode == 3   ? tc::launch2<a, b, 4, -1, 0, true>(g, grid_sms(ctx), s)       \
         : g.mode == 1 ? tc::launch2<a, b, 4>(g, grid_sms(ctx), s)                    \
                       : tc::launch2<a, b, 2>(g, grid_sms(ctx), s)
    if (A_MN && B_MN) GO2(true, true);
    if (A_MN && !B_MN) GO2(true, false);
    if (!A_MN && B_MN) GO2(false, true);
    GO2(false, false);
#undef GO2
  }
  const int TN = pick_tn(g, ctx->sm_count);
  YOTA_TRY(tc::make_maps(&mp, g, A_MN, B_MN, TN));
#define GO(TNv, a, b)                                                                          \
  return g.mode == 3   ? tc::launch<TNv, a, b, 4, true>(g, mp, grid_sms(ctx), s)               \
         : g.mode == 1 ? tc::launch<TNv, a, b, 4>(g, mp, grid_sms(ctx), s)                     \
                       : tc::launch<TNv, a, b, 2>(g, mp, grid_sms(ctx), s)
#define GOALL(TNv)                          \
  do {                                      \
    if (A_MN && B_MN) GO(TNv, true, true);  \
    if (A_MN && !B_MN) GO(TNv, true, false); \
    if (!A_MN && B_MN) GO(TNv, false, true); \
    GO(TNv, false, false);                  \
  } while (0)
  if (TN == 256) GOALL(256);
  if (TN == 128) GOALL(128);
  if (TN == 64) GOALL(64);
  if (eb == 4) {
    if (use_splitk4(g, ctx->sm_count)) {  // (B K-major: the forward and dh products of a recurrence)
      if (A_MN) return tc::launch<32, true, false, 4, false, -1, 0, 4>(g, mp, grid_sms(ctx), s);
      return tc::launch<32, false, false, 4, false, -1, 0, 4>(g, mp, grid_sms(ctx), s);
    }
#define GO32(a, b)                                                                      \
  return g.mode == 3 ? tc::launch<32, a, b, 4, true>(g, mp, grid_sms(ctx), s)           \
                     : tc::launch<32, a, b, 4>(g, mp, grid_sms(ctx), s)
    if (A_MN && B_MN) GO32(true, true);
    if (A_MN && !B_MN) GO32(true, false);
    if (!A_MN && B_MN) GO32(false, true);
    GO32(false, false);
#undef GO32
  }
  YOTA_FAIL(YOTA_E_ARG, "internal: no tile width for this GEMM");
#undef GOALL
#undef GO
}

// lo = x - tf32(x): the low part of a Float32 operand for the compensated (3xTF32) mode.  tf32(x)
// is x with the 13 low mantissa bits cleared -- what the tensor core makes of a Float32 it reads
// as TF32 -- so hi needs no array of its own (the GEMM reads x itself) and x = hi + lo exactly.
__global__ void split_tf32_kernel(float4* __restrict__ lo, const float4* __restrict__ src, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    float4 r;
    r.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
    r.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
    r.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
    r.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
    lo[i] = r;
  }
}
int launch_split_tf32(float* lo, const float* src, long long n, cudaStream_t s) {
  if (n == 0) return YOTA_OK;
  if ((n & 3) || ((reinterpret_cast<uintptr_t>(lo) | reinterpret_cast<uintptr_t>(src)) & 15))
    YOTA_FAIL(YOTA_E_SHAPE, "split_tf32: element count and pointers must be multiples of 4 / 16 bytes");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<float4*>(lo), reinterpret_cast<const float4*>(src), n / 4);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// Float32 -> BF16 (round to nearest even) shadow of a GEMM operand
__global__ void cast_bf16_kernel(__nv_bfloat162* __restrict__ dst, const float2* __restrict__ src, long long n2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = __ldg(src + i);
    dst[i] = __floats2bfloat162_rn(v.x, v.y);
  }
}
int launch_cast_bf16(void* dst, const float* src, long long n, cudaStream_t s) {
  if (n == 0) return YOTA_OK;
  if (n & 1) YOTA_FAIL(YOTA_E_SHAPE, "cast_bf16: odd element count");
  long long blocks = (n / 2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(static_cast<__nv_bfloat162*>(dst), reinterpret_cast<const float2*>(src), n / 2);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

}  // namespace yota
