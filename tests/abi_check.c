/* Prints the layout of every struct and the value of every enumerator / flag of
 * include/yota_cuda.h as `name value` lines.  tests/test_abi.py compares them with what the two
 * bindings assume: yota.jl_b200/ffi.py (ctypes Structures, opcode constants) and
 * yota.jl_b200/julia/YotaB200.jl (isbits structs and constant tables, parsed as text -- no
 * Julia toolchain exists in this image).  Built by __graft_entry__.build() with plain gcc: the
 * header must stay C. */
#include <stddef.h>
#include <stdio.h>

#include "../include/yota_cuda.h"

#define SZ(T) printf("sizeof." #T " %zu\n", sizeof(T))
#define OFF(T, f) printf("offsetof." #T "." #f " %zu\n", offsetof(T, f))
#define VAL(x) printf(#x " %lld\n", (long long)(x))

int main(void) {
  SZ(yota_tensor_desc);
  OFF(yota_tensor_desc, dtype);
  OFF(yota_tensor_desc, ndim);
  OFF(yota_tensor_desc, dims);
  OFF(yota_tensor_desc, kind);
  OFF(yota_tensor_desc, slot);
  OFF(yota_tensor_desc, cval);
  SZ(yota_op_desc);
  OFF(yota_op_desc, opcode);
  OFF(yota_op_desc, n_in);
  OFF(yota_op_desc, in);
  OFF(yota_op_desc, out);
  OFF(yota_op_desc, _pad);
  OFF(yota_op_desc, iattr);
  OFF(yota_op_desc, fattr);
  VAL(YOTA_ABI_VERSION);
  VAL(YOTA_MAX_DIMS);
  VAL(YOTA_MAX_OP_INPUTS);
  VAL(YOTA_N_IATTR);
  VAL(YOTA_OK); VAL(YOTA_E_CUDA); VAL(YOTA_E_NCCL); VAL(YOTA_E_SHAPE); VAL(YOTA_E_UNSUPPORTED_OP); VAL(YOTA_E_NOMEM); VAL(YOTA_E_ARG);
  VAL(YOTA_F32); VAL(YOTA_I64);
  VAL(YOTA_T_INPUT); VAL(YOTA_T_TEMP); VAL(YOTA_T_OUTPUT); VAL(YOTA_T_CONST); VAL(YOTA_T_SEED);
  VAL(YOTA_EW_COPY); VAL(YOTA_EW_ADD); VAL(YOTA_EW_SUB); VAL(YOTA_EW_MUL); VAL(YOTA_EW_DIV); VAL(YOTA_EW_NEG);
  VAL(YOTA_EW_TANH); VAL(YOTA_EW_EXP); VAL(YOTA_EW_LOG); VAL(YOTA_EW_SIN); VAL(YOTA_EW_COS); VAL(YOTA_EW_SQRT);
  VAL(YOTA_EW_RELU); VAL(YOTA_EW_SIGMOID); VAL(YOTA_EW_SQUARE); VAL(YOTA_EW_GT0); VAL(YOTA_EW_ADDC); VAL(YOTA_EW_MULC);
  VAL(YOTA_EW_RSUBC); VAL(YOTA_EW_RDIVC); VAL(YOTA_EW_POWC);
  VAL(YOTA_OP_SUM); VAL(YOTA_OP_MATMUL); VAL(YOTA_OP_GETINDEX); VAL(YOTA_OP_UNGETINDEX); VAL(YOTA_OP_LOGSOFTMAX);
  VAL(YOTA_OP_RESHAPE); VAL(YOTA_OP_TRANSPOSE); VAL(YOTA_OP_CAT); VAL(YOTA_OP_SLICE);
  VAL(YOTA_IDX_COLON); VAL(YOTA_IDX_INT); VAL(YOTA_IDX_RANGE); VAL(YOTA_IDX_VEC);
  VAL(YOTA_FLAG_NO_FUSE); VAL(YOTA_FLAG_NO_GRAPH); VAL(YOTA_FLAG_GEMM_FP32); VAL(YOTA_FLAG_GEMM_TF32);
  VAL(YOTA_FLAG_GEMM_BF16); VAL(YOTA_FLAG_NO_STATIC); VAL(YOTA_FLAG_GEMM_TF32X3);
  return 0;
}
