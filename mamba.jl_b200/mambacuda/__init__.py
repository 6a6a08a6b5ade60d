"""mambacuda — host side of the B200-native batched MCMC engine behind Mamba.jl's sampler API.

`Engine` is the thin binding of the C ABI (include/mambacuda.h); `mambacuda.api` mirrors the
reference's Julia interface for the hot path (Model templates, AMWG/Slice/RWM/NUTS/HMC/AMM sampler
constructors, setsamplers!, mcmc, gelmandiag, summarystats) on top of it.
"""
from ._lib import MambaCudaError, lib  # noqa: F401
from .engine import Engine  # noqa: F401
