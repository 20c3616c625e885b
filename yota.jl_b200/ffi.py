"""ctypes binding of libyota_cuda.so (include/yota_cuda.h).  The shared library is the
product; this file only marshals pointers and sizes.  There is no fallback: if the library
is missing or a call fails, `YotaError` is raised."""
from __future__ import annotations

import ctypes as C
import os

from . import ir

_HERE = os.path.dirname(os.path.abspath(__file__))
# YOTA_LIB: load another build of the same ABI (A/B of kernel changes inside one GPU job)
LIB_PATH = os.environ.get("YOTA_LIB") or os.path.join(_HERE, "lib", "libyota_cuda.so")


class YotaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libyota_cuda: {msg} (status {code})")
        self.code = code


class TensorDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("ndim", C.c_int32), ("dims", C.c_int64 * ir.MAX_DIMS),
                ("kind", C.c_int32), ("slot", C.c_int32), ("cval", C.c_double)]


class OpDesc(C.Structure):
    _fields_ = [("opcode", C.c_int32), ("n_in", C.c_int32), ("in_", C.c_int32 * ir.MAX_OP_INPUTS),
                ("out", C.c_int32), ("_pad", C.c_int32), ("iattr", C.c_int64 * ir.N_IATTR),
                ("fattr", C.c_double * 2)]


# every symbol include/yota_cuda.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "yota_abi_version": (_i, []),
    "yota_last_error": (C.c_char_p, []),
    "yota_init": (_i, [_i, _pp]),
    "yota_shutdown": (_i, [_vp]),
    "yota_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "yota_alloc": (_i, [_vp, _sz, _pp]),
    "yota_free": (_i, [_vp, _vp]),
    "yota_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "yota_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "yota_host_alloc": (_i, [_vp, _sz, _pp]),
    "yota_host_free": (_i, [_vp, _vp]),
    "yota_h2d_async": (_i, [_vp, _vp, _vp, _sz]),
    "yota_d2h_async": (_i, [_vp, _vp, _vp, _sz]),
    "yota_stream_fence": (_i, [_vp, _i, _i]),
    "yota_memset": (_i, [_vp, _vp, _i, _sz]),
    "yota_sync": (_i, [_vp]),
    "yota_program_build": (_i, [_vp, C.POINTER(OpDesc), _i64, C.POINTER(TensorDesc), _i64, C.c_uint32, _pp]),
    "yota_program_run": (_i, [_vp, _pp, _i64, _pp, _i64, _f]),
    "yota_program_destroy": (_i, [_vp]),
    "yota_program_describe": (_i64, [_vp, C.c_char_p, _i64]),
    "yota_program_dump_ops": (_i64, [_vp, C.c_char_p, _i64]),
    "yota_program_stats": (_i, [_vp] + [C.POINTER(_i64)] * 5),
    "yota_event_create": (_i, [_vp, _pp]),
    "yota_event_record": (_i, [_vp, _vp]),
    "yota_event_elapsed_ms": (_i, [_vp, _vp, _vp, C.POINTER(_f)]),
    "yota_event_destroy": (_i, [_vp, _vp]),
    "yota_program_profile": (_i, [_vp, _pp, _i64, _pp, _i64, _f, C.POINTER(_f), _i64, C.c_char_p, _i64]),
    "yota_program_timeline": (_i, [_vp, _pp, _i64, _pp, _i64, _f, _i, C.POINTER(_f), _i64, C.c_char_p, _i64]),
    "yota_gemm": (_i, [_vp, _i, _i, _i, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i]),
    "yota_scatter_add_cols": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _i64]),
    "yota_scatter_add_cols_sharded": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64]),
    "yota_gather_cols": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _i64]),
    "yota_update_sgd": (_i, [_vp, _vp, _vp, _i64, _f]),
    "yota_update_adam": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i]),
    "yota_program_set_update_adam": (_i, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i64, _f, _f, _f, _f]),
    "yota_comm_unique_id": (_i, [_vp]),
    "yota_comm_init": (_i, [_vp, _i, _i, _vp]),
    "yota_comm_destroy": (_i, [_vp]),
    "yota_program_set_allreduce": (_i, [_vp, C.POINTER(C.c_int32), _i64]),
    "yota_allreduce_f32": (_i, [_vp, _vp, _i64]),
    "yota_program_set_update": (_i, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i64, _f]),
    "yota_comm_symm_available": (_i, [_vp]),
    "yota_comm_symm_alloc": (_i, [_vp, _sz, _pp]),
    "yota_comm_symm_free": (_i, [_vp, _vp]),
}

FLAG_NO_FUSE, FLAG_NO_GRAPH, FLAG_GEMM_TF32, FLAG_GEMM_BF16, FLAG_NO_STATIC = 0x1, 0x2, 0x10, 0x20, 0x40
FLAG_GEMM_TF32X3 = 0x80

_lib = None


def lib():
    """Load libyota_cuda.so (built in-tree by `__graft_entry__.build()`); fail loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YotaError(-6, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.yota_abi_version() != 1:
            raise YotaError(-6, "ABI version mismatch between ffi.py and libyota_cuda.so")
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise YotaError(code, lib().yota_last_error().decode("utf-8", "replace"))


def pack_program(prog: ir.IRProgram):
    """IRProgram -> (OpDesc[], TensorDesc[])"""
    T = (TensorDesc * len(prog.tensors))()
    for t, d in zip(prog.tensors, T):
        d.dtype, d.ndim, d.kind, d.slot, d.cval = t.dtype, len(t.dims), t.kind, t.slot, t.cval
        for i, n in enumerate(t.dims):
            d.dims[i] = n
    O = (OpDesc * max(1, len(prog.ops)))()
    for o, d in zip(prog.ops, O):
        d.opcode, d.n_in, d.out = o.opcode, len(o.ins), o.out
        for i, t in enumerate(o.ins):
            d.in_[i] = t
        for i, v in enumerate(o.iattr):
            d.iattr[i] = v
        for i, v in enumerate(o.fattr):
            d.fattr[i] = v
    return O, T
