"""yota_b200: B200-native execution backend for the gradient tape of dfdx/Yota.jl.

Same surface as the reference for the hot path -- `grad`, `gradtape`, `compile`, `reset`,
`update` -- executed by hand-written sm_100a kernels in `lib/libyota_cuda.so` behind the C
ABI of `include/yota_cuda.h`.  See DESIGN.md and INTEGRATION.md.
"""
from . import jl  # noqa: F401
from .api import GRAD_CACHE, CompiledGrad, compile, grad, grad_compile, gradtape, reset, update, update_adam  # noqa: F401
from .device import B200Array, context, cu  # noqa: F401
from .eager import play, rrule  # noqa: F401
from .ffi import YotaError  # noqa: F401
from .rules import NoRuleError  # noqa: F401
from .tape import AVal, NoTangent, ZeroTangent, colon  # noqa: F401
