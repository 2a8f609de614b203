"""snout_b200 -- B200-native IQ -> packet receive engine behind Snout's receiver interfaces.

Only the hot path of nislab/snout lives here: captured complex-float IQ -> decoded BLE
advertising / IEEE 802.15.4 frames.  The DSP is hand-written CUDA for sm_100a in
snout_b200/csrc (built into snout_b200/lib/libsnoutrx.so); Python reaches it through the C ABI
of include/snoutrx.h via ctypes (snout_b200._abi).  There is no CPU fallback.
"""
from . import chanplan  # noqa: F401

__all__ = ["chanplan", "RxEngine"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "RxEngine":
        from .engine import RxEngine
        return RxEngine
    raise AttributeError(name)
