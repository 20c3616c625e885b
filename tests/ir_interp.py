"""TEST-ONLY NumPy interpreter of the op program (include/yota_cuda.h vocabulary).

Lets the CPU test-suite check the tape -> op-program lowering against the oracle without
a GPU.  It is not part of the product and is never imported by the package."""
import numpy as np

from oracle import yota_oracle as O


def _pad(a, nd):
    return a.reshape(a.shape + (1,) * (nd - a.ndim), order="F") if a.ndim < nd else a


def run(prog, inputs, seed=1.0, dtype=np.float32):
    """inputs: list indexed by in slot.  Returns list of outputs indexed by out slot."""
    import yota_b200  # noqa: F401
    from yota_b200 import ir

    t = np.dtype(dtype).type
    vals = {}
    for T in prog.tensors:
        if T.kind == ir.T_INPUT:
            a = np.asarray(inputs[T.slot])
            a = a.astype(dtype) if a.dtype.kind == "f" else a
            vals[T.id] = np.asfortranarray(a) if a.ndim else a  # (asfortranarray promotes 0-d to 1-d)
            assert tuple(vals[T.id].shape) == T.dims, (vals[T.id].shape, T.dims)
        elif T.kind == ir.T_CONST:
            vals[T.id] = np.asarray(t(T.cval))
        elif T.kind == ir.T_SEED:
            vals[T.id] = np.asarray(t(seed))
    un = {ir.EW_NEG: lambda x: -x, ir.EW_TANH: np.tanh, ir.EW_EXP: np.exp, ir.EW_LOG: np.log, ir.EW_SIN: np.sin,
          ir.EW_COS: np.cos, ir.EW_SQRT: np.sqrt, ir.EW_RELU: lambda x: np.maximum(x, 0),
          ir.EW_SIGMOID: lambda x: 1 / (1 + np.exp(-x)), ir.EW_SQUARE: lambda x: x * x,
          ir.EW_GT0: lambda x: (x > 0).astype(x.dtype), ir.EW_COPY: lambda x: x}
    bi = {ir.EW_ADD: np.add, ir.EW_SUB: np.subtract, ir.EW_MUL: np.multiply, ir.EW_DIV: np.divide}
    for o in prog.ops:
        od = prog.tensors[o.out].dims
        ins = [vals[i] for i in o.ins]
        c = t(o.fattr[0]) if o.fattr else None
        if o.opcode in un:
            y = un[o.opcode](ins[0])
        elif o.opcode in bi:
            nd = max(ins[0].ndim, ins[1].ndim)
            y = bi[o.opcode](_pad(ins[0], nd), _pad(ins[1], nd))
        elif o.opcode == ir.EW_ADDC:
            y = ins[0] + c
        elif o.opcode == ir.EW_MULC:
            y = ins[0] * c
        elif o.opcode == ir.EW_RSUBC:
            y = c - ins[0]
        elif o.opcode == ir.EW_RDIVC:
            y = c / ins[0]
        elif o.opcode == ir.EW_POWC:
            y = ins[0] ** c
        elif o.opcode == ir.OP_SUM:
            x = ins[0]
            ax = tuple(d for d in range(x.ndim) if o.iattr[0] >> d & 1)
            y = x.sum(axis=ax, keepdims=True) * t(o.fattr[0])
        elif o.opcode == ir.OP_MATMUL:
            A, B = (a.reshape(-1, 1) if a.ndim == 1 else a for a in ins)
            y = (A.T if o.iattr[0] else A) @ (B.T if o.iattr[1] else B)
        elif o.opcode in (ir.OP_GETINDEX, ir.OP_UNGETINDEX):
            n = o.iattr[0]
            I = []
            for d in range(n):
                k, a, b = o.iattr[1 + 3 * d:4 + 3 * d]
                I.append(O.colon if k == ir.IDX_COLON else int(a) if k == ir.IDX_INT else
                         O.JlRange(a, b) if k == ir.IDX_RANGE else ins[1 + a])
            if o.opcode == ir.OP_GETINDEX:
                y = np.asarray(O.jl_getindex(ins[0], tuple(I)))
            else:
                y = O.jl_ungetindex(od, dtype, ins[0], tuple(I))
        elif o.opcode == ir.OP_LOGSOFTMAX:
            y = O.rrule_logsoftmax(ins[0])[0]
        elif o.opcode == ir.OP_RESHAPE:
            y = ins[0]
        elif o.opcode == ir.OP_TRANSPOSE:
            y = ins[0].reshape(1, -1) if ins[0].ndim == 1 and len(od) == 2 else ins[0].T
        elif o.opcode == ir.OP_CAT:
            dim = o.iattr[0]
            nd = max(dim, max(a.ndim for a in ins))
            y = np.concatenate([_pad(a, nd) for a in ins], axis=dim - 1)
        elif o.opcode == 200:  # internal zero-copy concatenation of the planner (members side by side)
            y = np.concatenate([vals[m] for m in o.members], axis=1)
        elif o.opcode == ir.OP_SLICE:
            dim, a, b = o.iattr[:3]
            x = _pad(ins[0], max(dim, ins[0].ndim))
            sl = [slice(None)] * x.ndim
            sl[dim - 1] = slice(a - 1, b)
            y = x[tuple(sl)]
        else:
            raise NotImplementedError(o.opcode)
        y = np.asarray(y, dtype=dtype)
        if o.opcode in un or o.opcode in bi or o.opcode in (ir.EW_ADDC, ir.EW_MULC, ir.EW_RSUBC, ir.EW_RDIVC, ir.EW_POWC):
            y = np.broadcast_to(_pad(y, len(od)), od)
        y = np.asarray(y).reshape(od, order="F")
        vals[o.out] = np.asfortranarray(y) if y.ndim > 1 else y
    outs = [None] * prog.n_outputs
    for T in prog.tensors:
        if T.kind == ir.T_OUTPUT:
            outs[T.slot] = vals[T.id]
    return outs
