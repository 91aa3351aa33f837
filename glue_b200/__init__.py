r"""
glue_b200 -- B200-native (sm_100a) implementation of SCGLUE's per-step training
hot path behind the reference's ``scglue.models`` API.

Python / PyTorch host code calls hand-written CUDA kernels through the C ABI of
``libglue_b200.so`` (``include/glue_b200.h``).  There is no CPU fallback.
"""

__version__ = "0.1.0"

import os as _os

# A training step runs as ~10 parallel branches (encoders, graph branch, one per decoder call, the
# discriminator pass) beside the input prefetch copy.  The driver default of 8 hardware work queues
# makes branches share queues (false dependencies: the H2D prefetch stalls behind the step).  Read
# when the CUDA context is created, so it has to be in the environment before the first CUDA call;
# an explicit user setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import num  # noqa: F401
