"""verifiable_mpc_b200 -- B200-native MSM / generator-fold engine behind the verifiable_mpc prover API.

Host code is plain Python over a ctypes C ABI (include/vmsm.h -> libvmsm.so, hand-written CUDA for sm_100a).
No PyTorch, no Triton, no CPU fallback.  See DESIGN.md.
"""
from ._lib import VmsmError  # noqa: F401
from .engine import Context, DevicePoints, DeviceScalars, default_context  # noqa: F401

__version__ = "0.1.0"
