// Strict-FP32 GEMM (FFMA pipe): the exactness variant of the matmul rrule
// (ChainRules rrule(*, A, B): C = A*B, dA = Ȳ*B', dB = A'*Ȳ -- dispatched at
// /root/reference/src/grad.jl:147 and forced by map(unthunk, .) at src/grad.jl:235;
// SURVEY.md §8(a) row 10).  Column-major operands, any of NN / NT / TN / TT through element
// strides, optional accumulate (C += ...) which absorbs the tape's broadcast(+, dW, old)
// (src/grad.jl:91).  128x128x16 CTA tile, 8x8 register tile per thread, double-buffered
// shared memory.  Deterministic: fixed k order per output element.
#include "common.h"

namespace yota {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

// AM: A is contiguous along m (non-transposed column-major); else contiguous along k.
// BNc: B is contiguous along n (transposed B); else contiguous along k (non-transposed).
template <bool AM, bool BNc>
__global__ void __launch_bounds__(GT, 2) gemm_ffma_kernel(long long M, long long N, long long K,
                                                        const float* __restrict__ A, long long sam, long long sak,
                                                        const float* __restrict__ B, long long sbk, long long sbn,
                                                        float* __restrict__ C, long long ldc, int accumulate) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM, n0 = (long long)blockIdx.y * BN;
  const int tx = t & 15, ty = t >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto load_tile = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int m, k;
      if (AM) {
        m = t & 127;
        k = (t >> 7) + 2 * i;
      } else {
        k = t & 15;
        m = (t >> 4) + 16 * i;
      }
      const long long gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < M && gk < K) ? __ldg(A + gm * sam + gk * sak) : 0.f;
      int n, kb;
      if (BNc) {
        n = t & 127;
        kb = (t >> 7) + 2 * i;
      } else {
        kb = t & 15;
        n = (t >> 4) + 16 * i;
      }
      const long long gn = n0 + n, gkb = k0 + kb;
      rb[i] = (gn < N && gkb < K) ? __ldg(B + gkb * sbk + gn * sbn) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int m, k;
      if (AM) {
        m = t & 127;
        k = (t >> 7) + 2 * i;
      } else {
        k = t & 15;
        m = (t >> 4) + 16 * i;
      }
      As[buf][k][m] = ra[i];
      int n, kb;
      if (BNc) {
        n = t & 127;
        kb = (t >> 7) + 2 * i;
      } else {
        kb = t & 15;
        n = (t >> 4) + 16 * i;
      }
      Bs[buf][kb][n] = rb[i];
    }
  };

  const long long nk = (K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (long long kt = 0; kt < nk; kt++) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: thread owns rows {tx*4..+3, 64+tx*4..+3} x cols {ty*4..+3, 64+ty*4..+3}
  const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int jh = 0; jh < 2; jh++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const long long n = n0 + jh * 64 + ty * 4 + j;
      if (n >= N) continue;
#pragma unroll
      for (int ih = 0; ih < 2; ih++) {
        const long long m = m0 + ih * 64 + tx * 4;
        float* dst = C + m + n * ldc;
        float v[4] = {acc[ih * 4 + 0][jh * 4 + j], acc[ih * 4 + 1][jh * 4 + j], acc[ih * 4 + 2][jh * 4 + j],
                      acc[ih * 4 + 3][jh * 4 + j]};
        if (vec_ok && m + 3 < M) {
          if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dst);
            v[0] += o.x;
            v[1] += o.y;
            v[2] += o.z;
            v[3] += o.w;
          }
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (m + q < M) dst[q] = accumulate ? dst[q] + v[q] : v[q];
        }
      }
    }
}

// Latency kernel for tiny problems (C1: 784-128-10 MLP at batch 64): one thread per output
// element, sequential k (deterministic), operands served from L1/L2.  The 128x128 tiling above
// would put such a GEMM on 1-7 CTAs and walk K serially on each.
__global__ void __launch_bounds__(256) gemm_small_kernel(long long M, long long N, long long K,
                                                         const float* __restrict__ A, long long sam, long long sak,
                                                         const float* __restrict__ B, long long sbk, long long sbn,
                                                         float* __restrict__ C, long long ldc, int accumulate) {
  const long long m = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  const long long n = (long long)blockIdx.y * 8 + (threadIdx.x >> 5);
  if (m >= M || n >= N) return;
  const float* a = A + m * sam;
  const float* b = B + n * sbn;
  float acc = 0.f;
  long long k = 0;
  for (; k + 4 <= K; k += 4) {
    const float a0 = __ldg(a + k * sak), a1 = __ldg(a + (k + 1) * sak), a2 = __ldg(a + (k + 2) * sak),
                a3 = __ldg(a + (k + 3) * sak);
    const float b0 = __ldg(b + k * sbk), b1 = __ldg(b + (k + 1) * sbk), b2 = __ldg(b + (k + 2) * sbk),
                b3 = __ldg(b + (k + 3) * sbk);
    acc = fmaf(a0, b0, acc);
    acc = fmaf(a1, b1, acc);
    acc = fmaf(a2, b2, acc);
    acc = fmaf(a3, b3, acc);
  }
  for (; k < K; k++) acc = fmaf(__ldg(a + k * sak), __ldg(b + k * sbk), acc);
  float* dst = C + m + n * ldc;
  *dst = accumulate ? *dst + acc : acc;
}

// Same tiny problems when K is long (C1: K = 784, 128): one WARP per output element, lanes
// stride K (coalesced when K is the contiguous axis), fixed xor-shuffle tree at the end --
// deterministic, and 32x less serial latency than one thread walking K.
__global__ void __launch_bounds__(256) gemm_small_warp_kernel(long long M, long long N, long long K,
                                                              const float* __restrict__ A, long long sam, long long sak,
                                                              const float* __restrict__ B, long long sbk, long long sbn,
                                                              float* __restrict__ C, long long ldc, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= M * N) return;
  const long long m = w % M, n = w / M;
  const float* a = A + m * sam;
  const float* b = B + n * sbn;
  float acc = 0.f;
  for (long long k = lane; k < K; k += 32) acc = fmaf(__ldg(a + k * sak), __ldg(b + k * sbk), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  if (lane == 0) {
    float* dst = C + m + n * ldc;
    *dst = accumulate ? *dst + acc : acc;
  }
}

// Tiny problems whose op(A) has M contiguous (C1 forward W1*x, dW1 = dz1*x'): lanes take 32
// consecutive rows (coalesced A, broadcast B), the 8 warps of the block split K, and warp 0 adds
// the 8 partial sums in warp order -- deterministic, and 8x less serial latency per output.
__global__ void __launch_bounds__(256) gemm_small_splitk_kernel(long long M, long long N, long long K,
                                                                const float* __restrict__ A, long long sak,
                                                                const float* __restrict__ B, long long sbk, long long sbn,
                                                                float* __restrict__ C, long long ldc, int accumulate) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long m = (long long)blockIdx.x * 32 + lane, n = blockIdx.y;
  const long long k0 = K * w / 8, k1 = K * (w + 1) / 8;
  float acc = 0.f;
  if (m < M) {
    const float* a = A + m;
    const float* b = B + n * sbn;
    long long k = k0;
    for (; k + 4 <= k1; k += 4) {
      const float a0 = __ldg(a + k * sak), a1 = __ldg(a + (k + 1) * sak), a2 = __ldg(a + (k + 2) * sak),
                  a3 = __ldg(a + (k + 3) * sak);
      const float b0 = __ldg(b + k * sbk), b1 = __ldg(b + (k + 1) * sbk), b2 = __ldg(b + (k + 2) * sbk),
                  b3 = __ldg(b + (k + 3) * sbk);
      acc = fmaf(a0, b0, acc);
      acc = fmaf(a1, b1, acc);
      acc = fmaf(a2, b2, acc);
      acc = fmaf(a3, b3, acc);
    }
    for (; k < k1; k++) acc = fmaf(__ldg(a + k * sak), __ldg(b + k * sbk), acc);
  }
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && m < M) {
    float t = part[0][lane];
#pragma unroll
    for (int i = 1; i < 8; i++) t = __fadd_rn(t, part[i][lane]);
    float* dst = C + m + n * ldc;
    *dst = accumulate ? *dst + t : t;
  }
}

int launch_gemm_ffma(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return YOTA_OK;
  // element strides: op(A)(m,k), op(B)(k,n)
  const long long sam = g.transA ? g.lda : 1, sak = g.transA ? 1 : g.lda;
  const long long sbk = g.transB ? g.ldb : 1, sbn = g.transB ? 1 : g.ldb;
  const bool tiny = g.M * g.N <= 128 * 1024 && g.M * g.N * g.K <= (64ll << 20);
  if (tiny && sam == 1 && g.K >= 32 && g.N <= 65535) {
    dim3 grid((unsigned)((g.M + 31) / 32), (unsigned)g.N);
    gemm_small_splitk_kernel<<<grid, 256, 0, s>>>(g.M, g.N, g.K, static_cast<const float*>(g.A), sak,
                                                  static_cast<const float*>(g.B), sbk, sbn, g.C, g.ldc, g.accumulate);
    YOTA_CUDA(cudaGetLastError());
    return YOTA_OK;
  }
  if (tiny && sak == 1 && sbk == 1 && g.K >= 128) {  // both operands K-contiguous: lanes stride K
    const long long warps = g.M * g.N;
    gemm_small_warp_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(g.M, g.N, g.K, static_cast<const float*>(g.A), sam,
                                                                       sak, static_cast<const float*>(g.B), sbk, sbn, g.C,
                                                                       g.ldc, g.accumulate);
    YOTA_CUDA(cudaGetLastError());
    return YOTA_OK;
  }
  if (g.M * g.N <= 128 * 1024 && g.M * g.N * g.K <= (64ll << 20) && (g.N + 7) / 8 <= 65535) {
    dim3 grid((unsigned)((g.M + 31) / 32), (unsigned)((g.N + 7) / 8));
    gemm_small_kernel<<<grid, 256, 0, s>>>(g.M, g.N, g.K, static_cast<const float*>(g.A), sam, sak,
                                           static_cast<const float*>(g.B), sbk, sbn, g.C, g.ldc, g.accumulate);
    YOTA_CUDA(cudaGetLastError());
    return YOTA_OK;
  }
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (unsigned)((g.N + BN - 1) / BN));
  if (grid.y > 65535) YOTA_FAIL(YOTA_E_SHAPE, "gemm: N too large for the FFMA kernel grid");
  const bool AM = !g.transA, BNc = g.transB != 0;
#define LAUNCH(a, b) \
  gemm_ffma_kernel<a, b><<<grid, GT, 0, s>>>(g.M, g.N, g.K, static_cast<const float*>(g.A), sam, sak,          \
                                             static_cast<const float*>(g.B), sbk, sbn, g.C, g.ldc, g.accumulate)
  if (AM && BNc)
    LAUNCH(true, true);
  else if (AM && !BNc)
    LAUNCH(true, false);
  else if (!AM && BNc)
    LAUNCH(false, true);
  else
    LAUNCH(false, false);
#undef LAUNCH
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

}  // namespace yota
