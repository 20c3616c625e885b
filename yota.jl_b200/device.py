"""Device context and the device array type (the `CuArray` of this backend).

`B200Array` is a column-major Float32 (or Int64 index) array living in HBM, owned through
`yota_alloc` / `yota_free`.  Passing B200Arrays to `grad` is what routes a call to the
backend, the way array-type dispatch routes the reference to CUDA.jl
(/root/reference/test/test_grad.jl:262-267, test/test_helpers.jl:74-86).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import ffi

_ctxs = {}


class Context:
    def __init__(self, device):
        self.device = device
        h = C.c_void_p()
        ffi.check(ffi.lib().yota_init(device, C.byref(h)))
        self.h = h
        sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        ffi.check(ffi.lib().yota_device_info(h, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
        self.sm_count, self.cc, self.total_mem = sm.value, (maj.value, mnr.value), mem.value

    def sync(self):
        ffi.check(ffi.lib().yota_sync(self.h))


def context(device=None) -> Context:
    """The per-process context; one process per GPU (LOCAL_RANK picks the device)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ctxs:
        _ctxs[device] = Context(device)
    return _ctxs[device]


_NP = {"f32": np.float32, "i64": np.int64}


class B200Array:
    """Column-major device array.  `dims` follow Julia's size(x)."""

    def __init__(self, dims, dtype="f32", ctx=None, ptr=None, owner=True):
        self.ctx = ctx or context()
        self.dims = tuple(int(d) for d in dims)
        self.dtype = dtype
        self.nbytes = int(np.prod(self.dims, dtype=np.int64)) * np.dtype(_NP[dtype]).itemsize if self.dims else np.dtype(_NP[dtype]).itemsize
        self._owner = owner
        if ptr is None:
            p = C.c_void_p()
            ffi.check(ffi.lib().yota_alloc(self.ctx.h, max(self.nbytes, 4), C.byref(p)))
            ptr = p.value
        self.ptr = ptr

    shape = property(lambda self: self.dims)
    size = shape

    @property
    def numel(self):
        return int(np.prod(self.dims, dtype=np.int64)) if self.dims else 1

    @classmethod
    def from_host(cls, a, ctx=None):
        a = np.asarray(a)
        dt = "i64" if a.dtype.kind in "iu" else "f32"
        h = np.asfortranarray(a, dtype=_NP[dt])
        out = cls(h.shape, dt, ctx)
        if h.size:
            ffi.check(ffi.lib().yota_h2d(out.ctx.h, out.ptr, h.ctypes.data, h.nbytes))
        return out

    def copy_from_host(self, a):
        h = np.asfortranarray(a, dtype=_NP[self.dtype])
        assert h.shape == self.dims
        ffi.check(ffi.lib().yota_h2d(self.ctx.h, self.ptr, h.ctypes.data, h.nbytes))

    def to_host(self):
        out = np.empty(self.dims, dtype=_NP[self.dtype], order="F")
        if out.size:
            ffi.check(ffi.lib().yota_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out if self.dims else out[()]

    def __del__(self):
        try:
            if self._owner and self.ptr and ffi._lib is not None:
                ffi._lib.yota_free(self.ctx.h, self.ptr)
        except Exception:
            pass
        self.ptr = None

    def __repr__(self):
        return f"B200Array{{{'Float32' if self.dtype == 'f32' else 'Int64'}}}{list(self.dims)}"


def cu(a, ctx=None):
    """Move a host array to the device (CUDA.jl's `cu`)."""
    return a if isinstance(a, B200Array) else B200Array.from_host(a, ctx)
