"""dicp_b200 — B200-native (sm_100a) implementation of dICP's differentiable-ICP inner loop.

Host side mirrors the reference's Python surface (`dICP.ICP.ICP`, `dICP.nn.nn`,
`dICP.loss.loss`); the compute runs in hand-written CUDA kernels behind the C ABI declared
in include/dicp_b200.h (libdicp_b200.so, loaded with ctypes).  There is no CPU fallback:
calling the solver without a CUDA device or without the built library raises.
"""
from .ICP import ICP          # noqa: F401
from .nn import nn            # noqa: F401
from .loss import loss        # noqa: F401

__version__ = "0.1.0"
