// Micro-program tables of the ahead-of-time specialised elementwise chains (shared by the
// region kernels in ew_kernels.cu and the GEMM epilogues in gemm_tc.cu).
#pragma once
#include <type_traits>
#include <utility>

#include "common.h"
#include "ew_ops.cuh"

namespace yota {

constexpr int CH_MAX_OPS = 48;

struct ChainProg {
  int V;
  int n_in, n_ops, n_out, n_slots;
  int in_cls[RG_MAX_IN], in_slot[RG_MAX_IN];
  int op[CH_MAX_OPS], dst[CH_MAX_OPS], a[CH_MAX_OPS], b[CH_MAX_OPS];
  int out_kind[RG_MAX_OUT], out_src[RG_MAX_OUT], out_acc[RG_MAX_OUT];
};

template <int ID>
struct ChainOf;

// compile-time loop: f(integral_constant<int, 0>{}), ..., f(integral_constant<int, N-1>{})
template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}


#include "ew_static_chains.inc"

}  // namespace yota
