"""dppo_b200: ctypes binding of libdppo_b200.so, the sm_100a DPPO update engine.

PyTorch tensors are only the device-memory container; every computation on the hot path is a
hand-written CUDA kernel behind the C ABI in include/dppo_b200.h.  There is no CPU fallback:
importing the engine without the built library raises.
"""
from ._lib import load_library, DppoConfig, LibraryMissing, GAE_REFERENCE, GAE_STANDARD  # noqa: F401
from .layout import ParamLayout  # noqa: F401
from .engine import DppoEngine  # noqa: F401
