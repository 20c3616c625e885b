"""ctypes loader of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(so)
        _lib.oracle_gather_dot_f64.restype = C.c_double
    return _lib


def scatter_add_cols(dy, idx, n_dst):
    """dx = zero(rows x n_dst); dx[:, idx[k]] += dy[:, k] sequentially (1-based idx)."""
    dy = np.asfortranarray(dy, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    rows, n_src = dy.shape
    dx = np.empty((rows, n_dst), dtype=np.float32, order="F")
    lib().oracle_scatter_add_cols_f32(dx.ctypes.data_as(C.c_void_p), C.c_int64(rows), C.c_int64(n_dst),
                                      dy.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), C.c_int64(n_src))
    return dx


def gather_cols(x, idx):
    x = np.asfortranarray(x, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    y = np.empty((x.shape[0], idx.size), dtype=np.float32, order="F")
    lib().oracle_gather_cols_f32(y.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.c_int64(x.shape[0]),
                                 idx.ctypes.data_as(C.c_void_p), C.c_int64(idx.size))
    return y


def gather_dot(E, V, idx):
    E = np.asfortranarray(E, dtype=np.float32)
    V = np.asfortranarray(V, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    return lib().oracle_gather_dot_f64(E.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), C.c_int64(E.shape[0]),
                                       idx.ctypes.data_as(C.c_void_p), C.c_int64(idx.size))
