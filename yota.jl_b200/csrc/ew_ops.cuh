// Scalar semantics of the elementwise op vocabulary (include/yota_cuda.h: YOTA_EW_*), shared
// by the interpreter, the precompiled chains and the generic N-d fallback.  IEEE
// round-to-nearest single ops, never contracted into FMAs, precise (<= 2 ulp) libdevice
// transcendentals: no --use_fast_math anywhere.
#pragma once
#include "../../include/yota_cuda.h"

namespace yota {

template <int V>
struct VecT;
template <>
struct VecT<1> {
  using type = float;
  static __device__ __forceinline__ float splat(float v) { return v; }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float hsum(float a) { return a; }
};
template <>
struct VecT<4> {
  using type = float4;
  static __device__ __forceinline__ float4 splat(float v) { return make_float4(v, v, v, v); }
  static __device__ __forceinline__ float4 add(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
  }
  static __device__ __forceinline__ float hsum(float4 a) {
    return __fadd_rn(__fadd_rn(a.x, a.y), __fadd_rn(a.z, a.w));
  }
};

__host__ __device__ __forceinline__ bool ew_is_binary(int op) { return op >= YOTA_EW_ADD && op <= YOTA_EW_DIV; }

template <int OP>
__device__ __forceinline__ float ew_scalar(float a, float b, float imm) {
  if constexpr (OP == YOTA_EW_COPY) return a;
  if constexpr (OP == YOTA_EW_ADD) return __fadd_rn(a, b);
  if constexpr (OP == YOTA_EW_SUB) return __fsub_rn(a, b);
  if constexpr (OP == YOTA_EW_MUL) return __fmul_rn(a, b);
  if constexpr (OP == YOTA_EW_DIV) return __fdiv_rn(a, b);
  if constexpr (OP == YOTA_EW_NEG) return -a;
  if constexpr (OP == YOTA_EW_TANH) return tanhf(a);
  if constexpr (OP == YOTA_EW_EXP) return expf(a);
  if constexpr (OP == YOTA_EW_LOG) return logf(a);
  if constexpr (OP == YOTA_EW_SIN) return sinf(a);
  if constexpr (OP == YOTA_EW_COS) return cosf(a);
  if constexpr (OP == YOTA_EW_SQRT) return __fsqrt_rn(a);
  if constexpr (OP == YOTA_EW_RELU) return fmaxf(a, 0.f);
  if constexpr (OP == YOTA_EW_SIGMOID) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-a)));
  if constexpr (OP == YOTA_EW_SQUARE) return __fmul_rn(a, a);
  if constexpr (OP == YOTA_EW_GT0) return a > 0.f ? 1.f : 0.f;
  if constexpr (OP == YOTA_EW_ADDC) return __fadd_rn(a, imm);
  if constexpr (OP == YOTA_EW_MULC) return __fmul_rn(a, imm);
  if constexpr (OP == YOTA_EW_RSUBC) return __fsub_rn(imm, a);
  if constexpr (OP == YOTA_EW_RDIVC) return __fdiv_rn(imm, a);
  if constexpr (OP == YOTA_EW_POWC) return powf(a, imm);
  return a;
}

template <int OP>
__device__ __forceinline__ float ew_v(float a, float b, float imm) {
  return ew_scalar<OP>(a, b, imm);
}
template <int OP>
__device__ __forceinline__ float4 ew_v(float4 a, float4 b, float imm) {
  return make_float4(ew_scalar<OP>(a.x, b.x, imm), ew_scalar<OP>(a.y, b.y, imm), ew_scalar<OP>(a.z, b.z, imm),
                     ew_scalar<OP>(a.w, b.w, imm));
}

// runtime-dispatched form (interpreter): one warp-uniform switch per micro-op
template <int V>
__device__ __forceinline__ typename VecT<V>::type ew_apply(int op, typename VecT<V>::type a,
                                                           typename VecT<V>::type b, float imm) {
  switch (op) {
#define YOTA_CASE(OP) \
  case OP:            \
    return ew_v<OP>(a, b, imm);
    YOTA_CASE(YOTA_EW_COPY)
    YOTA_CASE(YOTA_EW_ADD)
    YOTA_CASE(YOTA_EW_SUB)
    YOTA_CASE(YOTA_EW_MUL)
    YOTA_CASE(YOTA_EW_DIV)
    YOTA_CASE(YOTA_EW_NEG)
    YOTA_CASE(YOTA_EW_TANH)
    YOTA_CASE(YOTA_EW_EXP)
    YOTA_CASE(YOTA_EW_LOG)
    YOTA_CASE(YOTA_EW_SIN)
    YOTA_CASE(YOTA_EW_COS)
    YOTA_CASE(YOTA_EW_SQRT)
    YOTA_CASE(YOTA_EW_RELU)
    YOTA_CASE(YOTA_EW_SIGMOID)
    YOTA_CASE(YOTA_EW_SQUARE)
    YOTA_CASE(YOTA_EW_GT0)
    YOTA_CASE(YOTA_EW_ADDC)
    YOTA_CASE(YOTA_EW_MULC)
    YOTA_CASE(YOTA_EW_RSUBC)
    YOTA_CASE(YOTA_EW_RDIVC)
    YOTA_CASE(YOTA_EW_POWC)
#undef YOTA_CASE
    default:
      return a;
  }
}

}  // namespace yota
