// getindex (gather) and its pullback: a DETERMINISTIC, bit-exact scatter-add.
//
// Reference semantics: ungetindex! ->
// NNlib.scatter!(+, dx, dy, idx) at /root/reference/src/helpers.jl:7-11 and ChainRules'
// ∇getindex! both run   dx = zero(x); for k in eachindex(dy) (column-major): dx[I_k] += dy[k]
// i.e. per destination a LEFT FOLD IN SOURCE ORDER starting from +0.0f.  Atomics (what
// NNlibCUDA does) or tree reductions change the rounding, so instead:
//   1. stable LSD radix sort of (destination, source position) pairs  -> per destination the
//      source positions in increasing order (warp match-any multi-split, no atomics on data)
//   2. segment starts per destination
//   3. one thread per destination ELEMENT folds its contributions sequentially and writes the
//      element exactly once (so dx needs no memset and every byte of dx is written once).
// HBM traffic for dx[:, idx] += dy: read dy once (full 128-byte lines per gathered column),
// read idx, write dx once -- the algorithmic minimum of SURVEY §8(d) C5.
#include "common.h"

namespace yota {

// ------------------------------------------------------------------------------------------
// gather
// ------------------------------------------------------------------------------------------
// An out-of-range index is a BoundsError in the reference; here it is REPORTED (the next
// synchronising call fails with YOTA_E_SHAPE) and clamped so the kernel itself stays in bounds.
__device__ __forceinline__ long long checked_index(long long j1, long long n, long long* err) {
  long long j = j1 - 1;
  if (j < 0 || j >= n) {
    report_index_error(err, j1, n);
    j = j < 0 ? 0 : n - 1;
  }
  return j;
}

__global__ void gather_cols_v4_kernel(float4* __restrict__ y, const float4* __restrict__ x, int R4, long long n_dst,
                                      const int64_t* __restrict__ idx, long long n_src, long long* err) {
  const long long total = n_src * R4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / R4;
    const int r = (int)(i - k * R4);
    const long long j = checked_index(__ldg(idx + k), n_dst, err);
    y[i] = __ldg(x + j * R4 + r);
  }
}
__global__ void gather_cols_v1_kernel(float* __restrict__ y, const float* __restrict__ x, long long R, long long n_dst,
                                      const int64_t* __restrict__ idx, long long n_src, long long* err) {
  const long long total = n_src * R;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / R;
    const long long r = i - k * R;
    const long long j = checked_index(__ldg(idx + k), n_dst, err);
    y[i] = __ldg(x + j * R + r);
  }
}

static unsigned grid_for(long long total, int threads, int cap_blocks = 148 * 16) {
  long long b = (total + threads - 1) / threads;
  if (b > cap_blocks) b = cap_blocks;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int launch_gather_cols(float* y, const float* x, long long rows, long long n_dst, const int64_t* idx, long long n_src,
                       cudaStream_t s, long long* err) {
  if (rows * n_src == 0) return YOTA_OK;
  const bool v4 = (rows % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x)) % 16 == 0);
  if (v4)
    gather_cols_v4_kernel<<<grid_for(n_src * (rows / 4), 256), 256, 0, s>>>(
        reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), (int)(rows / 4), n_dst, idx, n_src, err);
  else
    gather_cols_v1_kernel<<<grid_for(n_src * rows, 256), 256, 0, s>>>(y, x, rows, n_dst, idx, n_src, err);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

struct GetIdxParams {
  int n;  // number of indices
  int kind[4];
  long long a[4];
  const int64_t* vec[4];
  long long xdim[4];     // extent of x along index d
  long long xstride[4];  // element stride of x along index d
  long long odim[4];     // extent of the (unsqueezed) output along index d (1 for INT)
  long long* err;        // out-of-range entries of an index vector are reported here
};

__global__ void getindex_generic_kernel(GetIdxParams p, float* __restrict__ y, const float* __restrict__ x,
                                        long long n_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out;
       i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, off = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (d < p.n) {
        const long long c = rem % p.odim[d];
        rem /= p.odim[d];
        long long xi;
        if (p.kind[d] == YOTA_IDX_COLON)
          xi = c;
        else if (p.kind[d] == YOTA_IDX_VEC) {
          xi = checked_index(__ldg(p.vec[d] + c), p.xdim[d], p.err);
        } else
          xi = p.a[d] - 1 + c;  // INT (c == 0) or RANGE
        off += xi * p.xstride[d];
      }
    }
    y[i] = __ldg(x + off);
  }
}

static void fill_getidx(const IndexSpec& sp, GetIdxParams& p) {
  p.n = sp.n;
  p.err = nullptr;
  long long st = 1;
  for (int d = 0; d < 4; d++) {
    p.kind[d] = 0;
    p.a[d] = 0;
    p.vec[d] = nullptr;
    p.xdim[d] = p.odim[d] = 1;
    p.xstride[d] = 0;
  }
  for (int d = 0; d < sp.n; d++) {
    p.kind[d] = sp.kind[d];
    p.a[d] = sp.a[d];
    p.vec[d] = sp.vec[d];
    p.xdim[d] = sp.xdims[d];
    p.xstride[d] = st;
    st *= sp.xdims[d];
    switch (sp.kind[d]) {
      case YOTA_IDX_COLON: p.odim[d] = sp.xdims[d]; break;
      case YOTA_IDX_INT: p.odim[d] = 1; break;
      case YOTA_IDX_RANGE: p.odim[d] = sp.b[d] - sp.a[d] + 1; break;
      default: p.odim[d] = sp.vlen[d]; break;
    }
  }
}

int launch_getindex(const IndexSpec& sp, float* y, const float* x, long long n_out, cudaStream_t s, long long* err) {
  if (n_out == 0) return YOTA_OK;
  // fast path: x[:, idx] on a matrix
  if (sp.n == 2 && sp.kind[0] == YOTA_IDX_COLON && sp.kind[1] == YOTA_IDX_VEC)
    return launch_gather_cols(y, x, sp.xdims[0], sp.xdims[1], sp.vec[1], sp.vlen[1], s, err);
  GetIdxParams p;
  fill_getidx(sp, p);
  p.err = err;
  getindex_generic_kernel<<<grid_for(n_out, 256), 256, 0, s>>>(p, y, x, n_out);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// in-place slice accumulation: the selection is walked in dy order (same index math as the
// gather), every selected element is touched exactly once
__global__ void slice_accum_kernel(GetIdxParams p, float* __restrict__ acc, const float* __restrict__ dy,
                                   long long n_sel, int scalar_dy) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sel;
       i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, off = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (d < p.n) {
        const long long c = rem % p.odim[d];
        rem /= p.odim[d];
        off += (p.kind[d] == YOTA_IDX_COLON ? c : p.a[d] - 1 + c) * p.xstride[d];
      }
    }
    const float v = __fadd_rn(0.f, dy[scalar_dy ? 0 : i]);
    acc[off] = __fadd_rn(v, acc[off]);
  }
}

int launch_slice_accum(const IndexSpec& sp, float* acc, const float* dy, bool scalar_dy, cudaStream_t s) {
  GetIdxParams p;
  fill_getidx(sp, p);
  long long n_sel = 1;
  for (int d = 0; d < sp.n; d++) n_sel *= p.odim[d];
  if (n_sel == 0) return YOTA_OK;
  slice_accum_kernel<<<grid_for(n_sel, 256), 256, 0, s>>>(p, acc, dy, n_sel, scalar_dy ? 1 : 0);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// exclusive scan of u32 (in place), any length; deterministic
// ------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ unsigned block_excl_scan_u32(unsigned v, unsigned* total, unsigned* sm /*>=8*/) {
  // exclusive scan of one value per thread across a 256-thread block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  unsigned wbase = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SCAN_T / 32; i++) {
    const unsigned t = sm[i];
    if (i < w) wbase += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return wbase + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) scan_tile_kernel(unsigned* __restrict__ data, long long n,
                                                           unsigned* __restrict__ bsum) {
  __shared__ unsigned sm[8];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS], t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    v[i] = (base + i < n) ? data[base + i] : 0u;
    t += v[i];
  }
  unsigned total;
  unsigned ex = block_excl_scan_u32(t, &total, sm);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < n) data[base + i] = ex;
    ex += v[i];
  }
  if (bsum && threadIdx.x == 0) bsum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T) scan_add_kernel(unsigned* __restrict__ data, long long n,
                                                          const unsigned* __restrict__ bsum) {
  const unsigned add = bsum[blockIdx.x];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) data[base + i] += add;
}

static size_t scan_ws_elems(long long n) {
  size_t tot = 0;
  while (n > SCAN_TILE) {
    n = (n + SCAN_TILE - 1) / SCAN_TILE;
    tot += (size_t)n;
  }
  return tot + 1;
}

static int exclusive_scan_u32(unsigned* data, long long n, unsigned* ws, cudaStream_t s) {
  if (n <= 0) return YOTA_OK;
  const long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb == 1) {
    scan_tile_kernel<<<1, SCAN_T, 0, s>>>(data, n, nullptr);
    YOTA_CUDA(cudaGetLastError());
    return YOTA_OK;
  }
  scan_tile_kernel<<<(unsigned)nb, SCAN_T, 0, s>>>(data, n, ws);
  YOTA_CUDA(cudaGetLastError());
  YOTA_TRY(exclusive_scan_u32(ws, nb, ws + nb, s));
  scan_add_kernel<<<(unsigned)nb, SCAN_T, 0, s>>>(data, n, ws);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (key = destination, value = source position)
// ------------------------------------------------------------------------------------------
constexpr int RS_T = 256, RS_WARPS = 8, RS_ROUNDS = 16, RS_TILE = RS_T * RS_ROUNDS;  // 4096 keys per CTA

// Destination range of a scatter-add: all of 0..n_total-1, or -- destination-sharded over GPUs --
// the slice [lo, lo + n_local) this rank owns.  Sources aimed elsewhere are skipped (they belong
// to another rank's slice); every rank scans all indices, nobody exchanges or reduces anything,
// and each destination's fold keeps its global source order: bit-identical to one GPU.
struct DstRange {
  long long n_total, lo, n_local;
};

__device__ __forceinline__ unsigned key_of_idx(const int64_t* __restrict__ idx, long long i, DstRange dr,
                                               long long* err = nullptr) {
  long long j = __ldg(idx + i) - 1;
  // out-of-range indices (a BoundsError in Julia, src/helpers.jl:7-11 -> NNlib.scatter!) are
  // reported by the first histogram pass (err != nullptr there) and go to an overflow bucket
  // that is never read, so the fold itself stays in bounds
  if (j < 0 || j >= dr.n_total) {
    report_index_error(err, j + 1, dr.n_total);
    return (unsigned)dr.n_local;
  }
  j -= dr.lo;
  return (j < 0 || j >= dr.n_local) ? (unsigned)dr.n_local : (unsigned)j;
}

// FROM_IDX: first pass -- keys are derived from the Int64 index vector on the fly and the value
// is the source position itself, so no separate (key, value) initialisation pass is needed.
template <int BITS, bool FROM_IDX>
__global__ void __launch_bounds__(RS_T) radix_hist_kernel(const unsigned* __restrict__ keys,
                                                          const int64_t* __restrict__ idx, DstRange n_dst,
                                                          long long n, int shift, unsigned* __restrict__ hist,
                                                          int nblocks, long long* err) {
  constexpr int ND = 1 << BITS;
  __shared__ unsigned sh[ND];
  for (int i = threadIdx.x; i < ND; i += RS_T) sh[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; r++) {
    const long long i = base + r * RS_T + threadIdx.x;
    if (i < n) {
      const unsigned k = FROM_IDX ? key_of_idx(idx, i, n_dst, err) : keys[i];
      atomicAdd(&sh[(k >> shift) & (ND - 1)], 1u);  // integer counts: order-independent
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ND; i += RS_T) hist[(long long)i * nblocks + blockIdx.x] = sh[i];
}

// Each warp owns a contiguous 512-key segment of the CTA's tile and walks it in order, 32
// keys at a time; ballots give the rank among equal digits inside a round.  Rank inside the
// CTA = digit's start in the tile + earlier warps' count + this warp's running count + rank in
// the round => a stable split.  The tile is first split in shared memory and then copied out
// digit run by digit run: consecutive threads write consecutive addresses (a direct scatter of
// 4-byte items costs a 32-byte sector each, which made this kernel L2-write bound at 60 us).
template <int BITS, bool FROM_IDX>
__global__ void __launch_bounds__(RS_T) radix_scatter_kernel(const unsigned* __restrict__ keys_in,
                                                             const unsigned* __restrict__ vals_in,
                                                             const int64_t* __restrict__ idx, DstRange n_dst,
                                                             unsigned* __restrict__ keys_out,
                                                             unsigned* __restrict__ vals_out, long long n, int shift,
                                                             const unsigned* __restrict__ hist, int nblocks) {
  constexpr int ND = 1 << BITS;
  static_assert(ND <= RS_T, "one thread per digit in the table pass");
  constexpr int SEG = RS_TILE / RS_WARPS, LD = 4;  // keys per warp; key loads kept in flight
  __shared__ unsigned sh[RS_WARPS * ND];  // [warp][digit] counts, then running tile-local offsets
  __shared__ unsigned gbase[ND];          // global position of tile-local position 0 of each digit run
  __shared__ unsigned xk[RS_TILE], xv[RS_TILE];
  __shared__ unsigned scan_sm[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned* mine = sh + w * ND;
  for (int i = threadIdx.x; i < RS_WARPS * ND; i += RS_T) sh[i] = 0;
  __syncthreads();
  const long long tbase = (long long)blockIdx.x * RS_TILE;
  const long long wbase = tbase + (long long)w * SEG;
  const unsigned lt = (1u << lane) - 1u;
  // Lanes holding the same digit, from BITS ballots (MATCH.ANY is microcoded and ~50x slower
  // than VOTE on this part).
  auto peers_of = [&](unsigned d, unsigned act) {
    unsigned peers = act;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
      const unsigned bit = (d >> b) & 1u;
      const unsigned v = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? v : ~v;
    }
    return peers;
  };
  auto load_key = [&](long long i) {
    return i < n ? (FROM_IDX ? key_of_idx(idx, i, n_dst) : keys_in[i]) : 0xffffffffu;
  };
  // pass 1: per-warp digit counts (keys are re-read in pass 2: 16 MB from L2 is cheaper than
  // 16 live registers per thread)
  for (int r0 = 0; r0 < RS_ROUNDS; r0 += LD) {
    unsigned k[LD];
#pragma unroll
    for (int q = 0; q < LD; q++) k[q] = load_key(wbase + (r0 + q) * 32 + lane);
#pragma unroll
    for (int q = 0; q < LD; q++) {
      const bool ok = wbase + (r0 + q) * 32 + lane < n;
      const unsigned d = (k[q] >> shift) & (ND - 1);
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      const unsigned peers = peers_of(d, act);
      if (ok && (peers & lt) == 0) mine[d] += __popc(peers);  // lowest lane of each digit group
      __syncwarp();
    }
  }
  __syncthreads();
  // per digit: start of its run in the tile (scan over digits of the tile totals), exclusive
  // prefix over warps, and where the run goes in the output (scan of hist)
  {
    const int d = threadIdx.x;
    unsigned tot = 0;
    if (d < ND) {
#pragma unroll
      for (int ww = 0; ww < RS_WARPS; ww++) tot += sh[ww * ND + d];
    }
    unsigned all;
    const unsigned lstart = block_excl_scan_u32(tot, &all, scan_sm);
    if (d < ND) {
      unsigned run = lstart;
#pragma unroll
      for (int ww = 0; ww < RS_WARPS; ww++) {
        const unsigned c = sh[ww * ND + d];
        sh[ww * ND + d] = run;
        run += c;
      }
      gbase[d] = hist[(long long)d * nblocks + blockIdx.x] - lstart;  // mod 2^32: added back below
    }
  }
  __syncthreads();
  // pass 2: stable split of the tile in shared memory
  for (int r0 = 0; r0 < RS_ROUNDS; r0 += LD) {
    unsigned k[LD], v[LD];
#pragma unroll
    for (int q = 0; q < LD; q++) {
      const long long i = wbase + (r0 + q) * 32 + lane;
      k[q] = load_key(i);
      v[q] = FROM_IDX ? (unsigned)i : (i < n ? vals_in[i] : 0u);
    }
#pragma unroll
    for (int q = 0; q < LD; q++) {
      const bool ok = wbase + (r0 + q) * 32 + lane < n;
      const unsigned d = (k[q] >> shift) & (ND - 1);
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      const unsigned peers = peers_of(d, act);
      const unsigned pos = mine[d] + __popc(peers & lt);
      if (ok) {
        xk[pos] = k[q];
        xv[pos] = v[q];
      }
      __syncwarp();
      if (ok && (peers & lt) == 0) mine[d] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  // copy out: tile-local position j of digit d goes to gbase[d] + j
  const long long left = n - tbase;
  const int nvalid = left < RS_TILE ? (int)left : RS_TILE;
#pragma unroll 4
  for (int j = threadIdx.x; j < nvalid; j += RS_T) {
    const unsigned key = xk[j];
    const unsigned pos = gbase[(key >> shift) & (ND - 1)] + (unsigned)j;
    keys_out[pos] = key;
    vals_out[pos] = xv[j];
  }
}

// seg[j] = first sorted position whose key >= j, for j in [0, n_dst]; keys sorted ascending
__global__ void seg_start_kernel(const unsigned* __restrict__ keys, long long n, unsigned* __restrict__ seg,
                                 long long n_dst) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += (long long)gridDim.x * blockDim.x) {
    const long long lo = (p == 0) ? 0 : (long long)keys[p - 1] + 1;
    long long hi = (p == n) ? n_dst : (long long)keys[p];
    if (hi > n_dst) hi = n_dst;
    for (long long j = lo; j <= hi; j++) seg[j] = (unsigned)p;
  }
}

struct SortPlan {
  int bits, passes;
  long long nblocks;
  size_t off_keys[2], off_vals[2], off_hist, off_seg, off_scan, total;
};

static SortPlan sort_plan(long long n_dst, long long n_src) {
  SortPlan sp{};
  int keybits = 1;
  while ((1ll << keybits) <= n_dst) keybits++;  // keys in [0, n_dst] (n_dst = overflow bucket)
  sp.passes = (keybits + 7) / 8;  // <= 8-bit digits: the per-warp rank tables stay small next to the tile
  sp.bits = (keybits + sp.passes - 1) / sp.passes;
  if (sp.bits < 4) sp.bits = 4;
  sp.nblocks = (n_src + RS_TILE - 1) / RS_TILE;
  if (sp.nblocks < 1) sp.nblocks = 1;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  for (int i = 0; i < 2; i++) {
    sp.off_keys[i] = o;
    o += al((size_t)n_src * 4 + 4);
    sp.off_vals[i] = o;
    o += al((size_t)n_src * 4 + 4);
  }
  sp.off_hist = o;
  const size_t hist_elems = ((size_t)1 << sp.bits) * (size_t)sp.nblocks;
  o += al(hist_elems * 4);
  sp.off_seg = o;
  o += al(((size_t)n_dst + 2) * 4);
  sp.off_scan = o;
  o += al(scan_ws_elems((long long)hist_elems) * 4 + 1024);
  sp.total = o;
  return sp;
}

template <int BITS, bool FROM_IDX>
static int radix_pass(const unsigned* kin, const unsigned* vin, const int64_t* idx, DstRange n_dst, unsigned* kout,
                      unsigned* vout, long long n, int shift, unsigned* hist, int nblocks, unsigned* scan_ws,
                      cudaStream_t s, long long* err) {
  constexpr int ND = 1 << BITS;
  radix_hist_kernel<BITS, FROM_IDX><<<nblocks, RS_T, 0, s>>>(kin, idx, n_dst, n, shift, hist, nblocks, err);
  YOTA_CUDA(cudaGetLastError());
  YOTA_TRY(exclusive_scan_u32(hist, (long long)ND * nblocks, scan_ws, s));
  radix_scatter_kernel<BITS, FROM_IDX><<<nblocks, RS_T, 0, s>>>(kin, vin, idx, n_dst, kout, vout, n, shift, hist, nblocks);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// Where a finished sort left its results in the workspace: *perm (source positions grouped by
// destination, ascending inside a group) and *seg (n_dst + 1 starts).
static void sort_results(const SortPlan& sp, void* ws, const unsigned** perm, const unsigned** seg) {
  char* base = static_cast<char*>(ws);
  const int cur = (sp.passes - 1) & 1;  // pass 0 writes buffer 0, every later pass flips
  *perm = reinterpret_cast<const unsigned*>(base + sp.off_vals[cur]);
  *seg = reinterpret_cast<const unsigned*>(base + sp.off_seg);
}

// Sort (idx -> dest key, position) and build segment starts.  Depends on the index vector
// only: a program may run it on a side stream long before the tangent it permutes exists.
static int sort_by_destination(const int64_t* idx, long long n_src, long long n_dst, void* ws, const unsigned** perm,
                               const unsigned** seg, cudaStream_t s, long long* err, long long dst_lo = 0,
                               long long n_total = -1) {
  const SortPlan sp = sort_plan(n_dst, n_src);
  const DstRange dr{n_total < 0 ? n_dst : n_total, dst_lo, n_dst};
  char* base = static_cast<char*>(ws);
  unsigned* keys[2] = {reinterpret_cast<unsigned*>(base + sp.off_keys[0]),
                       reinterpret_cast<unsigned*>(base + sp.off_keys[1])};
  unsigned* vals[2] = {reinterpret_cast<unsigned*>(base + sp.off_vals[0]),
                       reinterpret_cast<unsigned*>(base + sp.off_vals[1])};
  unsigned* hist = reinterpret_cast<unsigned*>(base + sp.off_hist);
  unsigned* segp = reinterpret_cast<unsigned*>(base + sp.off_seg);
  unsigned* scan_ws = reinterpret_cast<unsigned*>(base + sp.off_scan);
  int cur = 1;  // the first pass reads idx and writes buffer 0
  for (int pass = 0; pass < sp.passes; pass++) {
    const int shift = pass * sp.bits;
    int st;
    switch (sp.bits) {
#define RP(B)                                                                                                     \
  case B:                                                                                                         \
    st = pass == 0 ? radix_pass<B, true>(nullptr, nullptr, idx, dr, keys[0], vals[0], n_src, shift, hist,         \
                                         (int)sp.nblocks, scan_ws, s, err)                                        \
                   : radix_pass<B, false>(keys[cur], vals[cur], idx, dr, keys[cur ^ 1], vals[cur ^ 1], n_src,     \
                                          shift, hist, (int)sp.nblocks, scan_ws, s, nullptr);                     \
    break;
      RP(4) RP(5) RP(6) RP(7) RP(8)
#undef RP
      default:
        YOTA_FAIL(YOTA_E_ARG, "radix sort: unsupported digit width %d", sp.bits);
    }
    YOTA_TRY(st);
    cur = pass == 0 ? 0 : cur ^ 1;
  }
  seg_start_kernel<<<grid_for(n_src + 1, 256), 256, 0, s>>>(keys[cur], n_src, segp, n_dst);
  YOTA_CUDA(cudaGetLastError());
  sort_results(sp, ws, perm, seg);
  return YOTA_OK;
}

size_t scatter_cols_workspace_bytes(long long n_dst, long long n_src) { return sort_plan(n_dst, n_src).total; }

// ------------------------------------------------------------------------------------------
// ordered fold: dx[:, j] = ((0 + dy[:, k1]) + dy[:, k2]) + ...   k1 < k2 < ... with idx[k] == j
// ------------------------------------------------------------------------------------------
template <int V>
struct FV;
template <>
struct FV<4> {
  using T = float4;
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T add(T a, T b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
  }
  static __device__ __forceinline__ T mul(float s, T a) {
    return make_float4(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z), __fmul_rn(s, a.w));
  }
};
template <>
struct FV<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ T add(T a, T b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ T mul(float s, T a) { return __fmul_rn(s, a); }
};

template <int V, bool SCALED>
__global__ void __launch_bounds__(256) scatter_fold_cols_kernel(typename FV<V>::T* __restrict__ dx,
                                                                const typename FV<V>::T* __restrict__ dy, int RV,
                                                                long long n_dst, const unsigned* __restrict__ perm,
                                                                const unsigned* __restrict__ seg,
                                                                const float* __restrict__ scale) {
  using T = typename FV<V>::T;
  const float sc = SCALED ? __ldg(scale) : 1.f;
  auto ld = [&](long long off) {
    const T v = __ldg(dy + off);
    return SCALED ? FV<V>::mul(sc, v) : v;
  };
  // TPC threads cooperate on one destination column; a CTA holds 256 / TPC columns
  const int TPC = RV < 256 ? RV : 256;
  const int cpb = 256 / TPC;
  const int cs = threadIdx.x / TPC, rv0 = threadIdx.x % TPC;
  if (cs >= cpb) return;
  for (long long j = (long long)blockIdx.x * cpb + cs; j < n_dst; j += (long long)gridDim.x * cpb) {
    const unsigned p0 = __ldg(seg + j), p1 = __ldg(seg + j + 1);
    for (int rv = rv0; rv < RV; rv += TPC) {
      T acc = FV<V>::zero();
      unsigned p = p0;
      for (; p + 4 <= p1; p += 4) {  // 4 independent loads in flight, folded in source order
        const unsigned k0 = __ldg(perm + p), k1 = __ldg(perm + p + 1), k2 = __ldg(perm + p + 2), k3 = __ldg(perm + p + 3);
        const T v0 = ld((long long)k0 * RV + rv);
        const T v1 = ld((long long)k1 * RV + rv);
        const T v2 = ld((long long)k2 * RV + rv);
        const T v3 = ld((long long)k3 * RV + rv);
        acc = FV<V>::add(FV<V>::add(FV<V>::add(FV<V>::add(acc, v0), v1), v2), v3);
      }
      for (; p < p1; p++) acc = FV<V>::add(acc, ld((long long)__ldg(perm + p) * RV + rv));
      dx[j * RV + rv] = acc;
    }
  }
}

int launch_scatter_add_cols(float* dx, long long rows, long long n_dst, const float* dy, const int64_t* idx,
                            long long n_src, void* workspace, size_t ws_bytes, cudaStream_t s, const float* scale,
                            long long* err, long long dst_lo, long long n_total, int phase) {
  if (rows * n_dst == 0) return YOTA_OK;
  if (n_src >= (1ll << 32) - 1 || n_dst >= (1ll << 32) - 2)
    YOTA_FAIL(YOTA_E_SHAPE, "scatter-add: more than 2^32 sources/destinations are not supported");
  if (ws_bytes < scatter_cols_workspace_bytes(n_dst, n_src)) YOTA_FAIL(YOTA_E_ARG, "scatter-add: workspace too small");
  const unsigned *perm, *seg;
  if (phase == UNGET_FOLD)
    sort_results(sort_plan(n_dst, n_src), workspace, &perm, &seg);
  else
    YOTA_TRY(sort_by_destination(idx, n_src, n_dst, workspace, &perm, &seg, s, err, dst_lo, n_total));
  if (phase == UNGET_SORT) return YOTA_OK;
  const bool v4 = (rows % 4 == 0) && ((reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dy)) % 16 == 0);
  if (v4) {
    const int RV = (int)(rows / 4);
    const int TPC = RV < 256 ? RV : 256, cpb = 256 / TPC;
    const unsigned grid = grid_for((n_dst + cpb - 1) / cpb, 1, 148 * 32);
    if (scale)
      scatter_fold_cols_kernel<4, true><<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(dx),
                                                             reinterpret_cast<const float4*>(dy), RV, n_dst, perm, seg, scale);
    else
      scatter_fold_cols_kernel<4, false><<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(dx),
                                                              reinterpret_cast<const float4*>(dy), RV, n_dst, perm, seg, nullptr);
  } else {
    if (rows > 0x7fffffff) YOTA_FAIL(YOTA_E_SHAPE, "scatter-add: too many rows");
    const int RV = (int)rows;
    const int TPC = RV < 256 ? RV : 256, cpb = 256 / TPC;
    const unsigned grid = grid_for((n_dst + cpb - 1) / cpb, 1, 148 * 32);
    if (scale)
      scatter_fold_cols_kernel<1, true><<<grid, 256, 0, s>>>(dx, dy, RV, n_dst, perm, seg, scale);
    else
      scatter_fold_cols_kernel<1, false><<<grid, 256, 0, s>>>(dx, dy, RV, n_dst, perm, seg, nullptr);
  }
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// ------------------------------------------------------------------------------------------
// general ungetindex: any mix of : / int / range / (at most one) index vector
// ------------------------------------------------------------------------------------------
struct UngetParams {
  int nd;
  long long xdim[4];
  int kind[4];
  long long a[4], b[4];
  long long dystride[4];  // stride in dy of the coordinate along index d (0 for INT)
  int vec_dim;            // -1: none
  const unsigned* perm;
  const unsigned* seg;
};

__global__ void ungetindex_generic_kernel(UngetParams p, float* __restrict__ dx, const float* __restrict__ dy,
                                          long long n_x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_x; i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, off = 0, jv = 0;
    bool hit = true;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (d < p.nd) {
        const long long c = rem % p.xdim[d];
        rem /= p.xdim[d];
        const int kd = p.kind[d];
        if (kd == YOTA_IDX_COLON)
          off += c * p.dystride[d];
        else if (kd == YOTA_IDX_VEC)
          jv = c;
        else {
          hit = hit && (c + 1 >= p.a[d]) && (c + 1 <= p.b[d]);
          off += (c + 1 - p.a[d]) * p.dystride[d];
        }
      }
    }
    float acc = 0.f;
    if (hit) {
      if (p.vec_dim < 0) {
        acc = __fadd_rn(0.f, dy[off]);  // 0.0f + (-0.0f) = +0.0f, exactly like zero(x) .+= dy
      } else {
        const unsigned p0 = p.seg[jv], p1 = p.seg[jv + 1];
        const long long vs = p.dystride[p.vec_dim];
        for (unsigned q = p0; q < p1; q++) acc = __fadd_rn(acc, dy[off + (long long)p.perm[q] * vs]);
      }
    }
    dx[i] = acc;
  }
}

// kernels one ungetindex launches (the program's launch count is part of what the bench reports)
int ungetindex_launch_count(const IndexSpec& sp) {
  for (int d = 0; d < sp.n; d++)
    if (sp.kind[d] == YOTA_IDX_VEC) {
      const SortPlan pl = sort_plan(sp.xdims[d], sp.vlen[d]);
      long long n = ((long long)1 << pl.bits) * pl.nblocks;
      int scan = 0;  // exclusive_scan_u32: one kernel for a single tile, else tile + recursion + add
      for (;;) {
        const long long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
        if (nb == 1) {
          scan += 1;
          break;
        }
        scan += 2;
        n = nb;
      }
      return pl.passes * (2 + scan) + 2;  // per pass: histogram, scan, split; then segment starts and the fold
    }
  return 1;
}

size_t ungetindex_workspace_bytes(const IndexSpec& sp) {
  for (int d = 0; d < sp.n; d++)
    if (sp.kind[d] == YOTA_IDX_VEC) return scatter_cols_workspace_bytes(sp.xdims[d], sp.vlen[d]);
  return 0;
}

int launch_ungetindex(const IndexSpec& sp, float* dx, const float* dy, void* workspace, size_t ws_bytes, int sm_count,
                      cudaStream_t s, const float* scale, long long* err, int phase) {
  (void)sm_count;
  long long n_x = 1;
  for (int d = 0; d < sp.n; d++) n_x *= sp.xdims[d];
  if (n_x == 0) return YOTA_OK;
  int vec_dim = -1, nvec = 0;
  for (int d = 0; d < sp.n; d++)
    if (sp.kind[d] == YOTA_IDX_VEC) {
      vec_dim = d;
      nvec++;
    }
  if (nvec > 1) YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "ungetindex: more than one index vector is not supported");
  // fast path: dx[:, idx] += dy on a matrix
  if (sp.n == 2 && sp.kind[0] == YOTA_IDX_COLON && vec_dim == 1)
    return launch_scatter_add_cols(dx, sp.xdims[0], sp.xdims[1], dy, sp.vec[1], sp.vlen[1], workspace, ws_bytes, s,
                                   scale, err, 0, -1, phase);
  if (scale) YOTA_FAIL(YOTA_E_ARG, "internal: scaled source planned for a general ungetindex");
  UngetParams p{};
  p.nd = sp.n;
  p.vec_dim = vec_dim;
  long long st = 1;
  for (int d = 0; d < 4; d++) {
    p.xdim[d] = 1;
    p.kind[d] = YOTA_IDX_COLON;
    p.a[d] = p.b[d] = 1;
    p.dystride[d] = 0;
  }
  for (int d = 0; d < sp.n; d++) {
    p.xdim[d] = sp.xdims[d];
    p.kind[d] = sp.kind[d];
    p.a[d] = sp.a[d];
    p.b[d] = sp.b[d];
    long long ext;
    switch (sp.kind[d]) {
      case YOTA_IDX_COLON: ext = sp.xdims[d]; break;
      case YOTA_IDX_INT: ext = 1; break;
      case YOTA_IDX_RANGE: ext = sp.b[d] - sp.a[d] + 1; break;
      default: ext = sp.vlen[d]; break;
    }
    p.dystride[d] = (sp.kind[d] == YOTA_IDX_INT) ? 0 : st;
    st *= ext;
  }
  if (vec_dim >= 0) {
    if (ws_bytes < ungetindex_workspace_bytes(sp)) YOTA_FAIL(YOTA_E_ARG, "ungetindex: workspace too small");
    if (phase == UNGET_FOLD)
      sort_results(sort_plan(sp.xdims[vec_dim], sp.vlen[vec_dim]), workspace, &p.perm, &p.seg);
    else
      YOTA_TRY(sort_by_destination(sp.vec[vec_dim], sp.vlen[vec_dim], sp.xdims[vec_dim], workspace, &p.perm, &p.seg,
                                   s, err));
  }
  if (phase == UNGET_SORT) return YOTA_OK;
  ungetindex_generic_kernel<<<grid_for(n_x, 256), 256, 0, s>>>(p, dx, dy, n_x);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

}  // namespace yota
