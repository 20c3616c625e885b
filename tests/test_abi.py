"""The drop-in boundary without a Julia toolchain: the layout of include/yota_cuda.h as a C
compiler sees it (tests/abi_check.c) against what the two bindings assume -- the ctypes
Structures and constants of yota.jl_b200/ffi.py / ir.py, and the isbits structs and constant
tables of yota.jl_b200/julia/YotaB200.jl, parsed as text (the Julia file cannot be executed
here; this catches the mechanical half of a mismatch: field order, widths, padding, opcodes)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from yota_b200 import ffi, ir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "yota.jl_b200", "julia", "YotaB200.jl")


@pytest.fixture(scope="module")
def abi():
    exe = os.path.join(ROOT, "tests", "abi_check")
    src = exe + ".c"
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(ROOT, "include", "yota_cuda.h"))):
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", exe, src], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {k: int(v) for k, v in (l.split() for l in out.splitlines())}


def test_ctypes_structures_match_the_header(abi):
    for cname, S in (("yota_tensor_desc", ffi.TensorDesc), ("yota_op_desc", ffi.OpDesc)):
        assert C.sizeof(S) == abi[f"sizeof.{cname}"]
        for fname, _t in S._fields_:
            assert getattr(S, fname).offset == abi[f"offsetof.{cname}.{fname.rstrip('_')}"], (cname, fname)
    assert (ir.MAX_DIMS, ir.MAX_OP_INPUTS, ir.N_IATTR) == (abi["YOTA_MAX_DIMS"], abi["YOTA_MAX_OP_INPUTS"], abi["YOTA_N_IATTR"])
    for name, code in ir.OP_NAMES.items():
        pass
    for k, v in abi.items():
        if k.startswith("YOTA_EW_") or k.startswith("YOTA_OP_"):
            assert getattr(ir, k[5:]) == v, k
        if k.startswith("YOTA_IDX_"):
            assert getattr(ir, k[5:]) == v, k
        if k.startswith("YOTA_T_"):
            assert getattr(ir, k[5:]) == v, k
    assert (ffi.FLAG_NO_FUSE, ffi.FLAG_NO_GRAPH, ffi.FLAG_GEMM_TF32, ffi.FLAG_GEMM_BF16, ffi.FLAG_NO_STATIC,
            ffi.FLAG_GEMM_TF32X3) == tuple(abi[k] for k in ("YOTA_FLAG_NO_FUSE", "YOTA_FLAG_NO_GRAPH", "YOTA_FLAG_GEMM_TF32",
                                                            "YOTA_FLAG_GEMM_BF16", "YOTA_FLAG_NO_STATIC", "YOTA_FLAG_GEMM_TF32X3"))


JL_SIZES = {"Int32": 4, "Int64": 8, "Float64": 8, "Float32": 4, "UInt32": 4}


def _jl_struct_layout(text, name):
    """(field, offset) list and total size of an isbits Julia struct, by C layout rules (which is
    what Julia uses for isbits structs passed through ccall)."""
    body = re.search(r"^struct " + name + r"\n(.*?)^end", text, re.S | re.M).group(1)
    off, align_max, out = 0, 1, []
    for line in body.strip().splitlines():
        fname, ftype = [s.strip() for s in line.split("::")]
        m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", ftype)
        n, base = (int(m.group(1)), m.group(2)) if m else (1, ftype)
        sz = JL_SIZES[base]
        off = (off + sz - 1) // sz * sz
        out.append((fname, off))
        off += n * sz
        align_max = max(align_max, sz)
    return out, (off + align_max - 1) // align_max * align_max


def test_julia_glue_structs_and_constants_match_the_header(abi):
    text = open(JL).read()
    for jl_name, cname in (("TensorDesc", "yota_tensor_desc"), ("OpDesc", "yota_op_desc")):
        fields, size = _jl_struct_layout(text, jl_name)
        assert size == abi[f"sizeof.{cname}"], (jl_name, size)
        for fname, off in fields:
            assert off == abi[f"offsetof.{cname}.{fname}"], (jl_name, fname, off)
    # const EW = (copy=1, ...), const OP = (sum=64, ...): every opcode of the header, same value
    for table, prefix in (("EW", "YOTA_EW_"), ("OP", "YOTA_OP_")):
        body = re.search(r"^const " + table + r" = \((.*?)\)$", text, re.S | re.M).group(1)
        got = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(\d+)", body)}
        want = {k[len(prefix):].lower(): v for k, v in abi.items() if k.startswith(prefix) and not k.endswith("_LAST")}
        assert got == want, (table, set(got.items()) ^ set(want.items()))
    m = re.search(r"const IDX_COLON, IDX_INT, IDX_RANGE, IDX_VEC = (\d), (\d), (\d), (\d)", text)
    assert tuple(map(int, m.groups())) == tuple(abi[k] for k in ("YOTA_IDX_COLON", "YOTA_IDX_INT", "YOTA_IDX_RANGE", "YOTA_IDX_VEC"))
    flags = dict(re.findall(r"const (FLAG_\w+)\s*=\s*UInt32\((0x[0-9a-fA-F]+)\)", text))
    for k, v in flags.items():
        assert int(v, 16) == abi["YOTA_" + k], k
    assert {"FLAG_GEMM_TF32", "FLAG_GEMM_BF16", "FLAG_GEMM_TF32X3"} <= set(flags)
    # every C entry point the glue ccalls exists in the header with that name
    hdr = open(os.path.join(ROOT, "include", "yota_cuda.h")).read()
    declared = set(re.findall(r"\b(yota_[a-z0-9_]+)\s*\(", hdr))
    called = set(re.findall(r"ccall\(\(:(yota_[a-z0-9_]+),", text))
    assert called and called <= declared, called - declared
