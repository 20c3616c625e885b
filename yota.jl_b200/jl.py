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


sum = _dispatch("sum")  # noqa: A001
mean = _dispatch("mean")
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
