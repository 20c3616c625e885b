// Fused elementwise regions: broadcast chains, their pullbacks, unbroadcast / full-sum sinks
// and tape gradient accumulation in ONE pass over HBM.
//
// Replaces, for device arrays, the Base.Broadcast copyto! loops + Base.mapreduce passes the
// reference's compiled tape runs one op at a time (SURVEY.md §8(a) rows 5, 7, 8, 11:
// /root/reference/src/rulesets.jl:27-29 materialize, src/grad.jl:91 broadcast(+, dx, old),
// ChainRules broadcasted / sum rules dispatched at src/grad.jl:147).
//
// A region is an SSA micro-program over a 2-d (R rows x C cols, column-major) iteration
// space.  There is no runtime code generation: the generic kernel is a warp-uniform
// micro-op interpreter (every lane executes the same opcode -> no divergence) whose
// register file lives in shared memory; hot chains get the same micro-program compiled
// ahead of time (ew_static.cuh) so that values stay in registers.  Both use the
// non-contracting __fmul_rn/__fadd_rn forms, so interpreter and static kernels -- and the
// reference's unfused op-by-op arithmetic -- round identically.
//
// Thread mapping: thread t owns V (4 or 1) consecutive rows for the whole kernel and walks
// columns, so ROWVEC operands (bias) are loaded once and row-sums (db = sum(Δ; dims=2))
// accumulate in a fixed order per thread: partial[col-chunk][row], reduced by a second
// fixed-order pass.  Full sums use a fixed shuffle tree + per-CTA partials.  Everything is
// deterministic run to run.
#include "common.h"
#include "ew_ops.cuh"

namespace yota {

template <int V>
__device__ __forceinline__ typename VecT<V>::type ld_vec(const float* p);
template <>
__device__ __forceinline__ float ld_vec<1>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float4 ld_vec<4>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <int V>
__device__ __forceinline__ void st_vec(float* p, typename VecT<V>::type v);
template <>
__device__ __forceinline__ void st_vec<1>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void st_vec<4>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}

__device__ __forceinline__ float block_sum_fixed(float t, float* red /* >= 8 floats */) {
  // fixed xor-shuffle tree inside the warp, then warps 0..7 in order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = t;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) s = __fadd_rn(s, red[i]);
  }
  __syncthreads();
  return s;
}

// column base of a gathered operand: x[:, idx[c]]; an out-of-range index (BoundsError in the
// reference) is reported through in.err and clamped so the load stays in bounds
__device__ __forceinline__ long long gather_col(const RegionIn& in, long long c) {
  const long long j1 = __ldg(in.idx + c);
  long long j = j1 - 1;
  if (j < 0 || j >= in.gather_n) {
    report_index_error(in.err, j1, in.gather_n);
    j = j < 0 ? 0 : in.gather_n - 1;
  }
  return j;
}

template <int V>
__global__ void __launch_bounds__(RG_THREADS) region_interp_kernel(const __grid_constant__ RegionParams p) {
  using vec = typename VecT<V>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[8];
  vec* S = reinterpret_cast<vec*>(smem_raw) + threadIdx.x;  // slot s lives at S[s * RG_THREADS]
  const int tid = threadIdx.x;
  const int rb = blockIdx.x % p.rowblocks;
  const int chunk = blockIdx.x / p.rowblocks;
  const int rgl = tid % p.TRG, cc = tid / p.TRG;
  const long long RGn = p.R / V;
  const long long rg = (long long)rb * RG_THREADS + rgl;
  const bool active = (cc < p.TCC) && (rg < RGn);
  const long long row = rg * V;
  const long long c_begin = (long long)chunk * p.cols_per_chunk;
  const long long c_end = min(p.C, c_begin + p.cols_per_chunk);

  // prologue: operands that do not change along the columns, and zeroed accumulators
  for (int k = 0; k < p.n_in; k++) {
    const RegionIn in = p.in[k];
    if (in.cls == CLS_ROWVEC) {
      if (active) S[in.slot * RG_THREADS] = ld_vec<V>(in.ptr + row);
    } else if (in.cls == CLS_SCALAR) {
      S[in.slot * RG_THREADS] = VecT<V>::splat(__ldg(in.ptr));
    }
  }
  for (int k = 0; k < p.n_out; k++)
    if (p.out[k].kind != OUT_STORE) S[p.out[k].acc * RG_THREADS] = VecT<V>::splat(0.f);

  if (active) {
    for (long long c = c_begin + cc; c < c_end; c += p.TCC) {
      const long long off = c * p.R + row;
      // issue all streaming loads before the first use
      vec pre[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < p.n_in && p.in[k].cls == CLS_FULL) pre[k] = ld_vec<V>(p.in[k].ptr + off);
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (k < p.n_in && p.in[k].cls == CLS_FULL) S[p.in[k].slot * RG_THREADS] = pre[k];
      for (int k = 0; k < p.n_in; k++) {
        const RegionIn in = p.in[k];
        if (in.cls == CLS_FULL) {
          if (k >= 4) S[in.slot * RG_THREADS] = ld_vec<V>(in.ptr + off);
        } else if (in.cls == CLS_COLVEC) {
          S[in.slot * RG_THREADS] = VecT<V>::splat(__ldg(in.ptr + c));
        } else if (in.cls == CLS_GATHER) {
          S[in.slot * RG_THREADS] = ld_vec<V>(in.ptr + gather_col(in, c) * p.R + row);
        }
      }
      vec acc = VecT<V>::splat(0.f);
      for (int i = 0; i < p.n_ops; i++) {
        const MicroOp m = p.ops[i];
        const vec a = (m.a == RG_ACC) ? acc : S[m.a * RG_THREADS];
        vec b = a;
        if (ew_is_binary(m.op)) b = (m.b == RG_ACC) ? acc : S[m.b * RG_THREADS];
        acc = ew_apply<V>(m.op, a, b, m.imm);
        if (m.dst != RG_NONE) S[m.dst * RG_THREADS] = acc;
      }
      for (int k = 0; k < p.n_out; k++) {
        const RegionOut o = p.out[k];
        const vec v = S[o.src * RG_THREADS];
        if (o.kind == OUT_STORE)
          st_vec<V>(o.ptr + off, v);
        else
          S[o.acc * RG_THREADS] = VecT<V>::add(S[o.acc * RG_THREADS], v);
      }
    }
  }
  // epilogue: sums
  for (int k = 0; k < p.n_out; k++) {
    const RegionOut o = p.out[k];
    if (o.kind == OUT_ROWSUM) {
      if (active) st_vec<V>(o.ptr + ((long long)chunk * p.TCC + cc) * p.R + row, S[o.acc * RG_THREADS]);
    } else if (o.kind == OUT_SCALARSUM) {
      const float t = block_sum_fixed(VecT<V>::hsum(S[o.acc * RG_THREADS]), red);
      if (tid == 0) o.ptr[blockIdx.x] = t;
    }
  }
}

}  // namespace yota

#include "ew_static.cuh"

namespace yota {

void region_geometry(RegionLaunch& L, int sm_count, int ctas_per_sm) {
  RegionParams& p = L.p;
  const long long RG = p.R / L.V;
  if (RG >= RG_THREADS) {
    p.TRG = RG_THREADS;
    p.TCC = 1;
    p.rowblocks = (int)((RG + RG_THREADS - 1) / RG_THREADS);
  } else {
    p.TRG = (int)RG;
    p.TCC = RG_THREADS / p.TRG;
    if (p.TCC > p.C) p.TCC = (int)(p.C > 0 ? p.C : 1);
    p.rowblocks = 1;
  }
  L.smem = (size_t)p.n_slots * RG_THREADS * sizeof(float) * L.V;
  int per_sm = ctas_per_sm;
  if (L.smem > 0) {
    int fit = (int)((200 * 1024) / L.smem);
    if (fit < per_sm) per_sm = fit < 1 ? 1 : fit;
  }
  const long long target = (long long)sm_count * per_sm;
  const long long col_tiles = (p.C + p.TCC - 1) / p.TCC;
  long long chunks = (target + p.rowblocks - 1) / p.rowblocks;
  if (chunks > col_tiles) chunks = col_tiles;
  if (chunks < 1) chunks = 1;
  const long long tiles_per_chunk = (col_tiles + chunks - 1) / chunks;
  p.cols_per_chunk = tiles_per_chunk * p.TCC;
  p.colchunks = (int)((col_tiles + tiles_per_chunk - 1) / tiles_per_chunk);
  L.grid = p.rowblocks * p.colchunks;
}

int launch_region(const RegionLaunch& L, cudaStream_t s) {
  if (L.static_id >= 0) return launch_static_chain(L, s);
  static bool attr4 = false, attr1 = false;  // opt in to > 48 KiB dynamic smem once (not a stream op)
  if (L.V == 4) {
    if (!attr4) {
      YOTA_CUDA(cudaFuncSetAttribute(region_interp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr4 = true;
    }
    region_interp_kernel<4><<<L.grid, RG_THREADS, L.smem, s>>>(L.p);
  } else {
    if (!attr1) {
      YOTA_CUDA(cudaFuncSetAttribute(region_interp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr1 = true;
    }
    region_interp_kernel<1><<<L.grid, RG_THREADS, L.smem, s>>>(L.p);
  }
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// ---- finishers ---------------------------------------------------------------------------
// 32 rows x 8 partial-lanes per CTA: lane j folds partials j, j+8, ... in order, then the 8
// lane sums are folded in order 0..7 -- a fixed order, 8x the parallelism of one thread per row.
__global__ void __launch_bounds__(256) rowsum_finish_kernel(float* __restrict__ out, const float* __restrict__ partial,
                                                            long long R, int P, float scale) {
  __shared__ float sm[8][33];
  const int rl = threadIdx.x & 31, j = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * 32 + rl;
  float s = 0.f;
  if (r < R)
    for (int q = j; q < P; q += 8) s = __fadd_rn(s, __ldg(partial + (long long)q * R + r));
  sm[j][rl] = s;
  __syncthreads();
  if (j == 0 && r < R) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) t = __fadd_rn(t, sm[q][rl]);
    out[r] = scale == 1.f ? t : __fmul_rn(t, scale);
  }
}

int launch_rowsum_finish(float* out, const float* partial, long long R, int P, float scale, cudaStream_t s) {
  rowsum_finish_kernel<<<(unsigned)((R + 31) / 32), 256, 0, s>>>(out, partial, R, P, scale);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

__global__ void __launch_bounds__(256) scalar_finish_kernel(float* out, const float* __restrict__ partial, int n,
                                                            float scale) {
  __shared__ float red[8];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) t = __fadd_rn(t, partial[i]);
  const float s = block_sum_fixed(t, red);
  if (threadIdx.x == 0) out[0] = scale == 1.f ? s : __fmul_rn(s, scale);
}

int launch_scalar_finish(float* out, const float* partial, int n, float scale, cudaStream_t s) {
  scalar_finish_kernel<<<1, 256, 0, s>>>(out, partial, n, scale);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// one warp per column; lanes stride the rows, fixed shuffle tree
__global__ void __launch_bounds__(256) colsum_kernel(float* __restrict__ out, const float* __restrict__ x, long long R,
                                                     long long C, float scale) {
  const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  const float* col = x + c * R;
  float t = 0.f;
  for (long long r = lane; r < R; r += 32) t = __fadd_rn(t, __ldg(col + r));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, o));
  if (lane == 0) out[c] = scale == 1.f ? t : __fmul_rn(t, scale);
}

int launch_colsum(float* out, const float* x, long long R, long long C, float scale, cudaStream_t s) {
  colsum_kernel<<<(unsigned)((C + 7) / 8), 256, 0, s>>>(out, x, R, C, scale);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// ---- generic N-d fallbacks ---------------------------------------------------------------
__global__ void generic_ew_kernel(GenericEw g, float* __restrict__ out, const float* __restrict__ a,
                                  const float* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, oa = 0, ob = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      if (d < g.nd) {
        const long long q = rem % g.dims[d];
        rem /= g.dims[d];
        oa += q * g.sa[d];
        ob += q * g.sb[d];
      }
    }
    const float av = a[oa];
    const float bv = b ? b[ob] : av;
    out[i] = ew_apply<1>(g.opcode, av, bv, g.imm);
  }
}

int launch_generic_ew(const GenericEw& g, float* out, const float* a, const float* b, cudaStream_t s) {
  long long n = 1;
  for (int d = 0; d < g.nd; d++) n *= g.dims[d];
  if (n == 0) return YOTA_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  generic_ew_kernel<<<(unsigned)blocks, 256, 0, s>>>(g, out, a, b, n);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// one thread per output element, sequential fold over the reduced dims (column-major order)
__global__ void generic_sum_kernel(GenericSum g, float* __restrict__ out, const float* __restrict__ x, long long n_out,
                                   long long n_red) {
  long long stride[4], kd[4], rd[4], ks[4], rs[4];
  int nk = 0, nr = 0;
  long long st = 1;
  for (int d = 0; d < g.nd; d++) {
    stride[d] = st;
    st *= g.dims[d];
    if (g.mask >> d & 1) {
      rd[nr] = g.dims[d];
      rs[nr++] = stride[d];
    } else {
      kd[nk] = g.dims[d];
      ks[nk++] = stride[d];
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out;
       i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, base = 0;
    for (int d = 0; d < nk; d++) {
      base += (rem % kd[d]) * ks[d];
      rem /= kd[d];
    }
    float sacc = 0.f;
    for (long long j = 0; j < n_red; j++) {
      long long r2 = j, o = base;
      for (int d = 0; d < nr; d++) {
        o += (r2 % rd[d]) * rs[d];
        r2 /= rd[d];
      }
      sacc = __fadd_rn(sacc, x[o]);
    }
    out[i] = g.scale == 1.f ? sacc : __fmul_rn(sacc, g.scale);
  }
}

int launch_generic_sum(const GenericSum& g, float* out, const float* x, cudaStream_t s) {
  long long n_out = 1, n_red = 1;
  for (int d = 0; d < g.nd; d++) (g.mask >> d & 1 ? n_red : n_out) *= g.dims[d];
  long long blocks = (n_out + 127) / 128;
  if (blocks > 148 * 16) blocks = 148 * 16;
  generic_sum_kernel<<<(unsigned)blocks, 128, 0, s>>>(g, out, x, n_out, n_red);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

__global__ void fill_const_kernel(float* out, float v) { out[0] = v; }
int launch_fill_const(float* out, float v, cudaStream_t s) {
  fill_const_kernel<<<1, 1, 0, s>>>(out, v);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

}  // namespace yota
