"""Julia-flavoured function namespace for writing loss functions against traced arrays.

`jl.sum(x)`, `jl.tanh(x)` (= `tanh.(x)`), `x @ W` (= `x * W`), `x * y` (= `x .* y`),
`x[jl.colon, idx]` ... dispatch on the array type through `__jl__`, the Python stand-in for
Julia's multiple dispatch on `CuArray` (/root/reference/test/test_grad.jl:262-267): the same
function body runs on this package's traced device arrays and on any other array type that
implements the protocol.  Indices are Julia indices: 1-based, ranges inclusive.
"""
from __future__ import annotations

from .tape import Colon, JlRange, colon  # noqa: F401


def range_(a, b):
    """a:b"""
    return JlRange(a, b)


def _dispatch(name):
    def f(x, *args, **kw):
        for a in (x,) + args:
            if hasattr(a, "__jl__"):
                return a.__jl__(name, x, *args, **kw)
        raise TypeError(f"{name}: no traced array among the arguments")

    f.__name__ = name
    return f


_sum = _dispatch("sum")
mean = _dispatch("mean")
reshape = _dispatch("reshape")


def sum(x, *args, **kw):  # noqa: A001
    """sum(x; dims) -- or sum(f, x): ChainRules' rule for the latter calls back into the AD through
    `rrule_via_ad(cfg, f, x_i)` (/root/reference/src/chainrules.jl:101-124), i.e. f's own tape is
    differentiated and spliced in.  Here f is traced in place (its statements become part of the
    parent tape -- the nested tape inlined as a sub-program) and applied elementwise."""
    if callable(x) and not hasattr(x, "__jl__"):
        return _sum(x(args[0]), **kw)
    return _sum(x, *args, **kw)


def map(f, x):  # noqa: A001
    """map(f, x) / broadcast(f, x) for a non-primitive f: the nested tape of f inlined (see sum)."""
    return f(x)
tanh = _dispatch("tanh")
exp = _dispatch("exp")
log = _dispatch("log")
sin = _dispatch("sin")
cos = _dispatch("cos")
abs2 = _dispatch("abs2")
sqrt = _dispatch("sqrt")
relu = _dispatch("relu")
sigmoid = _dispatch("sigmoid")
logsoftmax = _dispatch("logsoftmax")
getindex = _dispatch("getindex")
transpose = _dispatch("transpose")
adjoint = _dispatch("adjoint")
vcat = _dispatch("vcat")
hcat = _dispatch("hcat")
cat = _dispatch("cat")


# ---- composites over the device primitives (what NNlib / Flux provide to the reference's examples) ----
def logitcrossentropy(logits, y):
    """Flux.logitcrossentropy(ŷ, y) = mean(-sum(y .* logsoftmax(ŷ; dims=1); dims=1))
    (examples/flux/mlp.jl:81): written over the primitives, so its gradient is the tape's."""
    return mean(-_sum(y * logsoftmax(logits), dims=1))


def _conv_index(W, H, C, N, KW, KH):
    """1-based linear indices into x[W, H, C, N] of the im2col matrix P[(kw, kh, ci), (ow, oh, n)]
    of NNlib.conv (a true convolution: the kernel is flipped), stride 1, no padding."""
    import numpy as np

    OW, OH = W - KW + 1, H - KH + 1
    kw, kh, ci, ow, oh, n = np.meshgrid(np.arange(KW), np.arange(KH), np.arange(C), np.arange(OW), np.arange(OH),
                                        np.arange(N), indexing="ij")
    xi = (ow + (KW - 1 - kw)) + W * ((oh + (KH - 1 - kh)) + H * (ci + C * n))
    return (xi.reshape(-1, order="F") + 1).astype(np.int64), OW, OH


def conv(x, w):
    """NNlib.conv(x, w) for x[W, H, Cin, N], w[KW, KH, Cin, Cout] (stride 1, no padding; the layer of
    examples/flux/convnet.jl:53-66): im2col as a linear-index gather, one GEMM, and a gather that
    permutes the result to [OW, OH, Cout, N].  Every piece is a device primitive with a rule, so the
    pullback -- col2im as a bit-exact scatter-add, dW and dx as GEMMs -- comes from the tape."""
    import numpy as np

    W, H, C, N = x.shape
    KW, KH, C2, CO = w.shape
    assert C == C2, "DimensionMismatch: input channels"
    idx, OW, OH = _conv_index(W, H, C, N, KW, KH)
    P = reshape(x[idx], KW * KH * C, OW * OH * N)
    R = transpose(reshape(w, KW * KH * C, CO)) @ P  # [Cout, (ow, oh, n)]
    co, ow, oh, n = np.meshgrid(np.arange(CO), np.arange(OW), np.arange(OH), np.arange(N), indexing="ij")
    # y[ow, oh, co, n] = R[co, ow + OW * (oh + OH * n)]
    src = co + CO * (ow + OW * (oh + OH * n))
    perm = np.transpose(src, (1, 2, 0, 3)).reshape(-1, order="F") + 1
    return reshape(R[perm.astype(np.int64)], OW, OH, CO, N)


def maxpool2(x):
    """NNlib.maxpool(x, (2, 2)) for x[W, H, C, N] with even W, H: the four phase-shifted strided views
    as gathers, max(a, b) = a + relu(b - a) (ties route the gradient to the first candidate)."""
    import numpy as np

    W, H, C, N = x.shape
    assert W % 2 == 0 and H % 2 == 0
    ow, oh, c, n = np.meshgrid(np.arange(W // 2), np.arange(H // 2), np.arange(C), np.arange(N), indexing="ij")
    views = []
    for dh in (0, 1):
        for dw in (0, 1):
            lin = (2 * ow + dw) + W * ((2 * oh + dh) + H * (c + C * n))
            views.append(reshape(x[(lin.reshape(-1, order="F") + 1).astype(np.int64)], W // 2, H // 2, C, N))
    m = views[0]
    for v in views[1:]:
        m = m + relu(v - m)
    return m
