r"""
Data parallelism: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in CPU tests) for the plumbing.

The reference is single-device (``scglue/models/nn.py:180-220``); this module is new
design.  Cells shard across ranks (each rank runs a reference-identical step on its own
minibatch), the guidance graph and every parameter are replicated, and the only exchange
is one all-reduce (mean) of a flat gradient bucket per optimizer step -- ~12 MB for the
generator step at BASELINE config 1, latency-bound on NVSwitch, so it is sent as a single
message rather than per-parameter.
"""

import os
from typing import Iterable, Iterator, List

import torch
import torch.distributed as td


def is_distributed() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank() -> int:
    return td.get_rank() if td.is_available() and td.is_initialized() else 0


def world_size() -> int:
    return td.get_world_size() if td.is_available() and td.is_initialized() else 1


def init_from_env(backend: str = None) -> None:
    r"""Initialise the default process group from ``RANK`` / ``WORLD_SIZE`` / ``MASTER_*``
    (as set by ``torchrun``); binds the process to GPU ``LOCAL_RANK`` for NCCL."""
    if td.is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend=backend)


class GradBucket:
    r"""
    Flat fp32 buffer holding the gradients of a parameter list.  ``all_reduce_mean`` packs
    the current ``.grad`` tensors with one multi-tensor copy, averages the buffer across
    ranks with a single all-reduce and re-points every ``.grad`` at its slice, so the
    optimizer consumes the reduced values without a copy back.  Parameters whose gradient
    is ``None`` on this rank contribute zeros (ranks must agree on the message size).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]) -> None:
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("empty parameter list")
        device, total = self.params[0].device, sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views, offset = [], 0
        for p in self.params:
            self.views.append(self.flat[offset:offset + p.numel()].view_as(p))
            offset += p.numel()

    def all_reduce_mean(self) -> None:
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        missing = [v for v, p in zip(self.views, self.params) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if have:
            # a view that already is the .grad (previous step re-pointed it) needs no copy
            todo = [(v, g) for v, g in have if g.data_ptr() != v.data_ptr()]
            if todo:
                torch._foreach_copy_([v for v, _ in todo], [g for _, g in todo])
        if is_distributed():
            if td.get_backend() == "nccl":
                td.all_reduce(self.flat, op=td.ReduceOp.AVG)
            else:  # gloo has no AVG
                td.all_reduce(self.flat, op=td.ReduceOp.SUM)
                self.flat.div_(world_size())
        for v, p in zip(self.views, self.params):
            p.grad = v


class RankStridedBatches:
    r"""Wrap a batch sampler so that rank ``r`` of ``n`` yields batches ``r, r + n, ...``.
    The trailing ``len % n`` batches are dropped: every rank must take the same number of
    steps per epoch or the gradient all-reduce would dead-lock."""

    def __init__(self, batches, rank: int, world: int) -> None:
        self.batches, self.rank, self.world = batches, rank, world

    def __len__(self) -> int:
        return len(self.batches) // self.world

    def __iter__(self) -> Iterator[List[int]]:
        usable = len(self) * self.world
        for i, batch in enumerate(self.batches):
            if i >= usable:
                break
            if i % self.world == self.rank:
                yield batch


def mean_over_ranks(values: List[float], device: torch.device = None) -> List[float]:
    r"""Element-wise mean of a list of host floats over all ranks (identity when not
    distributed).  Training-loop decisions that every rank takes on its own copy of a
    metric -- learning-rate reduction, early stopping, the non-finite guard -- must see the
    same number everywhere, or one rank changes its learning rate / leaves the loop while the
    others wait in the next gradient all-reduce."""
    if not is_distributed() or not values:
        return list(values)
    on_gpu = td.get_backend() == "nccl"
    t = torch.tensor(values, dtype=torch.float64,
                     device=device if (on_gpu and device is not None) else ("cuda" if on_gpu else "cpu"))
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return (t / world_size()).tolist()


def broadcast_object(obj, src: int = 0):
    r"""``obj`` of rank ``src`` on every rank (identity when not distributed)."""
    if not is_distributed():
        return obj
    box = [obj]
    td.broadcast_object_list(box, src=src)
    return box[0]


def barrier() -> None:
    if is_distributed():
        td.barrier()
