r"""
glue_b200 -- B200-native (sm_100a) implementation of SCGLUE's per-step training
hot path behind the reference's ``scglue.models`` API.

Python / PyTorch host code calls hand-written CUDA kernels through the C ABI of
``libglue_b200.so`` (``include/glue_b200.h``).  There is no CPU fallback.
"""

__version__ = "0.1.0"

from . import num  # noqa: F401
