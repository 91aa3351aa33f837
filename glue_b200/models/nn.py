r"""
Network building blocks and device utilities (API of ``scglue/models/nn.py``).
"""

import os
import re

import numpy as np
import torch
from torch.nn.modules.batchnorm import _NormBase

from .. import dist as gdist
from .. import ops
from ..utils import config, logged


class GraphConv(torch.nn.Module):
    r"""
    Graph convolution (propagation only), same call signature as the reference
    (``scglue/models/nn.py:21-57``): ``res[t] = sum_{e: tidx_e = t} esgn_e enorm_e input[sidx_e]``.

    The edge list is converted once into target-sorted / source-sorted CSR (cached per
    edge tensors, see :func:`glue_b200.ops.csr_for`); forward and backward are one
    atomic-free warp-per-row gather kernel each, so results are reproducible run to run.
    """

    def forward(self, input: torch.Tensor, eidx: torch.Tensor, enorm: torch.Tensor,
                esgn: torch.Tensor) -> torch.Tensor:
        csr = ops.csr_for(eidx, enorm, esgn, input.shape[0])
        if config.PARTITION_GRAPH and gdist.is_distributed():
            # atlas scale: each rank aggregates a block of target rows, blocks are all-gathered
            part = ops.partition_for(csr, gdist.rank(), gdist.world_size())
            return ops.graph_conv_partitioned(input, part)
        return ops.graph_conv(input, csr)


def freeze_running_stats(m: torch.nn.Module) -> None:
    r"""Stop normalisation layers from updating running statistics (``nn.py:146-156``)."""
    if isinstance(m, _NormBase):
        m.eval()


def get_default_numpy_dtype(complex: bool = False) -> type:
    r"""numpy dtype matching the torch default dtype (``nn.py:159-173``)."""
    kind, bits = re.match(r"torch\.([a-z]+)(\d+)", str(torch.get_default_dtype())).groups()
    if complex:
        kind, bits = "complex", int(bits) * 2
    return getattr(np, f"{kind}{bits}")


def zero_nan_grad(grad: torch.Tensor) -> torch.Tensor:
    return grad.nan_to_num()


@logged
def autodevice() -> torch.device:
    r"""
    Computation device of this process: one process drives one GPU.  Under ``torchrun``
    the GPU is ``LOCAL_RANK``; otherwise the current CUDA device.  (The reference picks the
    GPU with most free memory, ``nn.py:180-220``; with one process per GPU that choice
    belongs to the launcher.)  ``config.CPU_ONLY`` yields a CPU device for host-side logic
    only -- the kernels themselves refuse CPU tensors.
    """
    if config.CPU_ONLY or not torch.cuda.is_available():
        if not config.CPU_ONLY:
            autodevice.logger.warning("No CUDA device visible: modules are built on CPU, "
                                      "but the sm_100a kernels cannot run there.")
        return torch.device("cpu")
    local_rank = os.environ.get("LOCAL_RANK")
    index = int(local_rank) if local_rank is not None else torch.cuda.current_device()
    return torch.device("cuda", index)
