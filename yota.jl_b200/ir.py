"""Flat op program (SSA tensor ids + ops) the gradient tape is lowered to.

Mirrors `include/yota_cuda.h` (yota_tensor_desc / yota_op_desc / yota_opcode) field for
field; `ffi.py` turns an `IRProgram` into the ctypes arrays `yota_program_build` takes.
The Julia glue (`julia/YotaB200.jl`) emits the same structs from a `Tape{GradCtx}`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

MAX_DIMS = 4
MAX_OP_INPUTS = 6
N_IATTR = 16

F32, I64 = 0, 1
T_INPUT, T_TEMP, T_OUTPUT, T_CONST, T_SEED = 0, 1, 2, 3, 4

# opcodes (include/yota_cuda.h: yota_opcode)
EW_COPY, EW_ADD, EW_SUB, EW_MUL, EW_DIV, EW_NEG, EW_TANH, EW_EXP, EW_LOG, EW_SIN, EW_COS = range(1, 12)
EW_SQRT, EW_RELU, EW_SIGMOID, EW_SQUARE, EW_GT0, EW_ADDC, EW_MULC, EW_RSUBC, EW_RDIVC, EW_POWC = range(12, 22)
OP_SUM, OP_MATMUL, OP_GETINDEX, OP_UNGETINDEX, OP_LOGSOFTMAX, OP_RESHAPE, OP_TRANSPOSE, OP_CAT, OP_SLICE = range(64, 73)

IDX_COLON, IDX_INT, IDX_RANGE, IDX_VEC = 0, 1, 2, 3

EW_NAMES = {
    EW_COPY: "copy", EW_ADD: "add", EW_SUB: "sub", EW_MUL: "mul", EW_DIV: "div", EW_NEG: "neg",
    EW_TANH: "tanh", EW_EXP: "exp", EW_LOG: "log", EW_SIN: "sin", EW_COS: "cos", EW_SQRT: "sqrt",
    EW_RELU: "relu", EW_SIGMOID: "sigmoid", EW_SQUARE: "square", EW_GT0: "gt0", EW_ADDC: "addc",
    EW_MULC: "mulc", EW_RSUBC: "rsubc", EW_RDIVC: "rdivc", EW_POWC: "powc",
}
OP_NAMES = dict(EW_NAMES)
OP_NAMES.update({OP_SUM: "sum", OP_MATMUL: "matmul", OP_GETINDEX: "getindex", OP_UNGETINDEX: "ungetindex",
                 OP_LOGSOFTMAX: "logsoftmax", OP_RESHAPE: "reshape", OP_TRANSPOSE: "transpose",
                 OP_CAT: "cat", OP_SLICE: "slice"})


class TId(int):
    """An op-program tensor id (distinct from plain ints so literal constants stay literal)."""

    def __repr__(self):
        return f"%{int(self)}"


@dataclass
class IRTensor:
    id: int
    dtype: int
    dims: tuple
    kind: int = T_TEMP
    slot: int = -1
    cval: float = 0.0

    @property
    def numel(self):
        n = 1
        for d in self.dims:
            n *= d
        return n


@dataclass
class IROp:
    opcode: int
    ins: tuple
    out: int
    iattr: tuple = ()
    fattr: tuple = ()


@dataclass
class IRProgram:
    tensors: list = field(default_factory=list)
    ops: list = field(default_factory=list)
    n_inputs: int = 0
    n_outputs: int = 0

    def tensor(self, dims, dtype=F32, kind=T_TEMP, slot=-1, cval=0.0):
        dims = tuple(int(d) for d in dims)
        if len(dims) > MAX_DIMS:
            raise ValueError(f"arrays of more than {MAX_DIMS} dims are not supported: {dims}")
        t = IRTensor(len(self.tensors), dtype, dims, kind, slot, float(cval))
        self.tensors.append(t)
        return TId(t.id)

    def const(self, v):
        return self.tensor((), F32, T_CONST, cval=v)

    def op(self, opcode, ins, out_dims, iattr=(), fattr=(), dtype=F32):
        assert len(ins) <= MAX_OP_INPUTS, "too many op inputs"
        out = self.tensor(out_dims, dtype)
        self.ops.append(IROp(opcode, tuple(ins), out, tuple(int(i) for i in iattr), tuple(float(f) for f in fattr)))
        return out

    def dims(self, tid):
        return self.tensors[tid].dims

    def mark_output(self, tid, slot):
        """Bind tensor `tid` to out_ptrs[slot].  Inputs/constants/aliases are copied into a
        fresh output tensor so every output has exactly one writer."""
        t = self.tensors[tid]
        if t.kind != T_TEMP or any(o.opcode == OP_RESHAPE and o.out == tid for o in self.ops):
            tid = self.op(EW_COPY, (tid,), t.dims)
            t = self.tensors[tid]
        t.kind, t.slot = T_OUTPUT, slot
        self.n_outputs = max(self.n_outputs, slot + 1)
        return tid

    def describe(self):
        lines = []
        for o in self.ops:
            ins = ", ".join(f"%{int(i)}{list(self.tensors[i].dims)}" for i in o.ins)
            lines.append(f"%{int(o.out)}{list(self.tensors[o.out].dims)} = {OP_NAMES[o.opcode]}({ins})"
                         + (f" i={list(o.iattr)}" if o.iattr else "") + (f" f={list(o.fattr)}" if o.fattr else ""))
        return "\n".join(lines)


def bcast_dims(*shapes):
    """Julia broadcast shape: align from dim 1, pad trailing 1s."""
    nd = max((len(s) for s in shapes), default=0)
    out = []
    for d in range(nd):
        n = 1
        for s in shapes:
            m = s[d] if d < len(s) else 1
            if m != 1:
                if n != 1 and n != m:
                    raise ValueError(f"DimensionMismatch: arrays could not be broadcast to a common size: {shapes}")
                n = m
        out.append(n)
    return tuple(out)
