// Diagnostic (1 GPU): how does a persistent tcgen05 GEMM behave while a few SMs are held by
// another kernel (standing in for the all-reduce CTAs)?  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/occupy_probe tests/occupy_probe.cu \
//        -Iinclude -Lyota.jl_b200/lib -lyota_cuda -Xlinker -rpath -Xlinker '$ORIGIN/../yota.jl_b200/lib'
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "yota_cuda.h"

__global__ void __launch_bounds__(1024, 1) occupier(long long cycles, int* sink) {
  const long long t0 = clock64();
  int x = 0;
  while (clock64() - t0 < cycles) x++;
  if (x == -1) *sink = x;
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(1024, 1) occupier2(long long cycles, int* sink) {
  const long long t0 = clock64();
  int x = 0;
  while (clock64() - t0 < cycles) x++;
  if (x == -1) *sink = x;
}

#define CK(x)                                                        \
  do {                                                               \
    int _r = (x);                                                    \
    if (_r != 0) {                                                   \
      printf("FAIL %s: %d %s\n", #x, _r, yota_last_error());         \
      return 1;                                                      \
    }                                                                \
  } while (0)

int main() {
  yota_ctx* ctx;
  CK(yota_init(0, &ctx));
  const long long M = 4096, N = 1024, K = 4096;
  void *A, *B, *C;
  CK(yota_alloc(ctx, M * K * 4, &A));
  CK(yota_alloc(ctx, K * N * 4, &B));
  CK(yota_alloc(ctx, M * N * 4, &C));
  CK(yota_memset(ctx, A, 0, M * K * 4));
  CK(yota_memset(ctx, B, 0, K * N * 4));
  cudaStream_t s2;
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  int* sink;
  cudaMalloc(&sink, 4);
  void *e0, *e1;
  CK(yota_event_create(ctx, &e0));
  CK(yota_event_create(ctx, &e1));
  auto gemm = [&](int reps) {
    for (int i = 0; i < reps; i++)
      if (yota_gemm(ctx, 1, 1, 0, M, N, K, (float*)A, K, (float*)B, K, (float*)C, M, 0)) return 1;
    return 0;
  };
  CK(gemm(3));
  CK(yota_sync(ctx));
  const char* grids[] = {"148", "132", "116"};
  for (const char* gsm : grids) {
    setenv("YOTA_GEMM_GRID_SMS", gsm, 1);
    for (int mode = 0; mode < 5; mode++) {
      // mode 0: alone; 1: 16 plain CTAs first; 2: 16 CTAs in clusters of 2 first; 3/4: same but GEMM first
      float best = 1e9f;
      for (int rep = 0; rep < 5; rep++) {
        CK(yota_sync(ctx));
        cudaDeviceSynchronize();
        const long long cyc = 600000;  // ~300 us
        if (mode == 1) occupier<<<16, 1024, 0, s2>>>(cyc, sink);
        if (mode == 2) occupier2<<<16, 1024, 0, s2>>>(cyc, sink);
        CK(yota_event_record(ctx, e0));
        CK(gemm(1));
        CK(yota_event_record(ctx, e1));
        if (mode == 3) occupier<<<16, 1024, 0, s2>>>(cyc, sink);
        if (mode == 4) occupier2<<<16, 1024, 0, s2>>>(cyc, sink);
        CK(yota_sync(ctx));
        cudaDeviceSynchronize();
        float ms;
        CK(yota_event_elapsed_ms(ctx, e0, e1, &ms));
        if (ms < best) best = ms;
      }
      const char* names[] = {"alone", "16 plain CTAs launched first", "16 CTAs in 2-clusters launched first",
                             "16 plain CTAs launched after", "16 CTAs in 2-clusters launched after"};
      printf("grid %s SMs, %-40s: dh GEMM 4096x1024x4096 tf32 %.1f us\n", gsm, names[mode], best * 1e3f);
    }
  }
  cudaError_t e = cudaGetLastError();
  printf("cuda: %s\n", cudaGetErrorString(e));
  return 0;
}
