// Column-wise logsoftmax (NNlib.logsoftmax(x; dims=1), the C3 loss head), transpose, strided
// copies for cat / slice pullbacks, and the update! axpy (/root/reference/src/update.jl:42-49).
#include "common.h"

namespace yota {

// One warp per column: max, then sum(exp(x - max)), then y = (x - max) - log(sum).  Lanes
// stride the rows; fixed xor-shuffle trees => deterministic.  The column is re-read from
// L1/L2 (16 KiB per 4096-row column), HBM sees one read and one write.
__global__ void __launch_bounds__(256) logsoftmax_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                         long long R, long long C) {
  const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  const float* xc = x + c * R;
  float* yc = y + c * R;
  float m = -INFINITY;
  for (long long r = lane; r < R; r += 32) m = fmaxf(m, xc[r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (long long r = lane; r < R; r += 32) s = __fadd_rn(s, expf(__fsub_rn(xc[r], m)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  const float ls = logf(s);
  for (long long r = lane; r < R; r += 32) yc[r] = __fsub_rn(__fsub_rn(xc[r], m), ls);
}

int launch_logsoftmax(float* y, const float* x, long long R, long long C, cudaStream_t s) {
  if (R * C == 0) return YOTA_OK;
  logsoftmax_kernel<<<(unsigned)((C + 7) / 8), 256, 0, s>>>(y, x, R, C);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// y (C x R) = x (R x C)', both column-major; 32x32 smem tiles, coalesced both ways
__global__ void __launch_bounds__(256) transpose_kernel(float* __restrict__ y, const float* __restrict__ x, long long R,
                                                        long long C) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + tx < R && c0 + j < C) tile[j][tx] = x[(r0 + tx) + (c0 + j) * R];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + tx < C && r0 + j < R) y[(c0 + tx) + (r0 + j) * C] = tile[tx][j];
}

int launch_transpose(float* y, const float* x, long long R, long long C, cudaStream_t s) {
  if (R * C == 0) return YOTA_OK;
  dim3 grid((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32));
  transpose_kernel<<<grid, 256, 0, s>>>(y, x, R, C);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

struct CopyParams {
  int nd;
  long long dims[4], ds[4], ss[4];
};
__global__ void copy_strided_kernel(CopyParams p, float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long rem = i, od = 0, os = 0;
#pragma unroll
    for (int d = 0; d < 4; d++)
      if (d < p.nd) {
        const long long c = rem % p.dims[d];
        rem /= p.dims[d];
        od += c * p.ds[d];
        os += c * p.ss[d];
      }
    dst[od] = src[os];
  }
}

int launch_copy_strided(float* dst, const float* src, int nd, const long long* dims, const long long* dst_stride,
                        const long long* src_stride, cudaStream_t s) {
  CopyParams p{};
  p.nd = nd;
  long long n = 1;
  for (int d = 0; d < nd; d++) {
    p.dims[d] = dims[d];
    p.ds[d] = dst_stride[d];
    p.ss[d] = src_stride[d];
    n *= dims[d];
  }
  if (n == 0) return YOTA_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  copy_strided_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, dst, src, n);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

__global__ void axpy_kernel(float* __restrict__ x, const float* __restrict__ g, long long n, float lr) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __fsub_rn(x[i], lr == 1.f ? g[i] : __fmul_rn(lr, g[i]));
}

int launch_axpy(float* x, const float* g, long long n, float lr, cudaStream_t s) {
  if (n == 0) return YOTA_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  axpy_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, g, n, lr);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

// Adam as Optimisers.jl applies it (examples/flux/mlp.jl:82, utils.jl:5 use Optimisers.Adam):
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  x -= m / (1 - b1^t) / (sqrt(v / (1 - b2^t)) + eps) * lr
// The step number lives on the device (t_dev: steps taken so far) so that a replayed CUDA graph
// advances it; t_host >= 1 is used by the stand-alone entry.  Non-contracted arithmetic: a Float32
// host evaluation of the same expression gives the same bits up to the libm powf / sqrtf.
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            const int* __restrict__ t_dev, int t_host) {
  const int t = t_dev ? *t_dev + 1 : t_host;
  const float c1 = __fsub_rn(1.f, powf(b1, (float)t)), c2 = __fsub_rn(1.f, powf(b2, (float)t));
  const float o1 = __fsub_rn(1.f, b1), o2 = __fsub_rn(1.f, b2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(o1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(o2, __fmul_rn(gi, gi)));
    m[i] = mi;
    v[i] = vi;
    const float step = __fmul_rn(__fdiv_rn(__fdiv_rn(mi, c1), __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, c2)), eps)), lr);
    x[i] = __fsub_rn(x[i], step);
  }
}
__global__ void tick_kernel(int* t) { *t += 1; }

int launch_adam(float* x, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                const int* t_dev, int t_host, cudaStream_t s) {
  if (n == 0) return YOTA_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, g, m, v, n, lr, b1, b2, eps, t_dev, t_host);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}
int launch_tick(int* t, cudaStream_t s) {
  tick_kernel<<<1, 1, 0, s>>>(t);
  YOTA_CUDA(cudaGetLastError());
  return YOTA_OK;
}

}  // namespace yota
