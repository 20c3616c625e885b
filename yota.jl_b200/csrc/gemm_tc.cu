// tcgen05 / TMEM tensor-core GEMM (placeholder until the kernel lands in this file).
#include "common.h"
namespace yota {
bool gemm_tc_eligible(const GemmArgs& g) { (void)g; return false; }
int launch_gemm_tc(yota_ctx* ctx, const GemmArgs& g, cudaStream_t s) {
  (void)ctx; (void)g; (void)s;
  YOTA_FAIL(YOTA_E_UNSUPPORTED_OP, "tensor-core GEMM not built");
}
}  // namespace yota
