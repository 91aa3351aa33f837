r"""
``FlatRMSprop``: RMSprop with every parameter, gradient and running average held in one flat
fp32 buffer each, updated by a single fused kernel per step.

The reference compiles its trainers with ``torch.optim.RMSprop`` defaults
(``scglue/models/scglue.py:889-942`` -> ``optim="RMSprop"``: ``alpha = 0.99``,
``eps = 1e-8``, no momentum, not centred, no weight decay).  Stock RMSprop walks the parameter
list with 6-7 multi-tensor launches per step; here

* parameters are re-pointed at slices of one flat buffer (``p.data`` stays a normal tensor
  view, so modules, ``state_dict`` and checkpoints are unaffected);
* gradients are packed into a flat buffer with one multi-tensor copy -- the same buffer the
  data-parallel all-reduce works on (:mod:`glue_b200.dist`), so gradient exchange needs no
  second copy;
* ``glue_rmsprop_flat`` applies the update in one pass (20 B of HBM traffic per parameter).

Semantics kept from ``torch.optim.RMSprop``: parameters whose ``.grad`` is ``None`` are
skipped entirely (their running average does not decay); ``param_groups[0]["lr"]`` may be
changed between steps (``ReduceLROnPlateau``); ``state_dict`` / ``load_state_dict`` use the
stock layout (``step``, ``square_avg`` per parameter).
"""

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as td

from . import _lib
from ._lib import check, current_stream, ptr


class FlatRMSprop(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-2, alpha: float = 0.99,
                 eps: float = 1e-8) -> None:
        params = list(params)
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        if isinstance(params[0], dict):
            raise ValueError("FlatRMSprop takes a single parameter group")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps))
        self._plist: List[torch.nn.Parameter] = self.param_groups[0]["params"]
        first = self._plist[0]
        if not first.is_cuda:
            raise RuntimeError("FlatRMSprop runs on CUDA parameters only (no CPU fallback)")
        for p in self._plist:
            if p.device != first.device or p.dtype != torch.float32:
                raise ValueError("FlatRMSprop needs fp32 parameters on one device")
        self.device = first.device
        # 16-byte aligned slices so that every contiguous run takes the vector path
        self._offsets, total = [], 0
        for p in self._plist:
            self._offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_sq = torch.zeros(total, dtype=torch.float32, device=self.device)
        self._pviews = [self._view(self.flat_param, i) for i in range(len(self._plist))]
        self._gviews = [self._view(self.flat_grad, i) for i in range(len(self._plist))]
        self._sviews = [self._view(self.flat_sq, i) for i in range(len(self._plist))]
        self._adopt_parameters()
        for p, sq in zip(self._plist, self._sviews):
            self.state[p] = {"step": torch.zeros((), dtype=torch.float32, device=self.device),
                             "square_avg": sq}
        self._packed: Optional[List[Tuple[int, int]]] = None

    # ------------------------------------------------------------------------------ plumbing
    def _view(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        p = self._plist[i]
        return flat[self._offsets[i]:self._offsets[i] + p.numel()].view_as(p)

    @torch.no_grad()
    def _adopt_parameters(self) -> None:
        r"""Point every ``p.data`` at its slice of the flat buffer (values preserved).  Called
        again whenever a parameter was re-allocated behind our back (``net.to(...)``)."""
        for p, view in zip(self._plist, self._pviews):
            if p.data_ptr() != view.data_ptr():
                view.copy_(p.data)
                p.data = view

    def _ranges(self, active: List[bool]) -> List[Tuple[int, int]]:
        r"""Contiguous flat ranges covering the active parameters."""
        ranges: List[Tuple[int, int]] = []
        for i, on in enumerate(active):
            if not on:
                continue
            start = self._offsets[i]
            stop = start + (self._plist[i].numel() + 3) // 4 * 4
            if ranges and ranges[-1][1] == start:
                ranges[-1] = (ranges[-1][0], stop)
            else:
                ranges.append((start, stop))
        return ranges

    @torch.no_grad()
    def pack_grads(self) -> List[Tuple[int, int]]:
        r"""Copy the current ``.grad`` tensors into the flat gradient buffer (one multi-tensor
        copy), re-point ``.grad`` at the slices and return the active flat ranges."""
        self._adopt_parameters()
        active = [p.grad is not None for p in self._plist]
        todo = [(v, p.grad) for v, p in zip(self._gviews, self._plist)
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [g for _, g in todo])
        for v, p, on in zip(self._gviews, self._plist, active):
            if on:
                p.grad = v
        self._packed = self._ranges(active)
        return self._packed

    @staticmethod
    def _all_reduce_avg(flat: torch.Tensor) -> None:
        if td.is_available() and td.is_initialized() and td.get_world_size() > 1 and flat.numel():
            if td.get_backend() == "nccl":
                td.all_reduce(flat, op=td.ReduceOp.AVG)
            else:
                td.all_reduce(flat, op=td.ReduceOp.SUM)
                flat.div_(td.get_world_size())

    @torch.no_grad()
    def all_reduce_subset(self, params) -> Optional[Tuple[int, int]]:
        r"""Exchange the gradients of ``params`` -- a contiguous run of this optimizer's parameters whose
        gradients are already final (the graph branch, back-propagated ahead of the rest of the objective) --
        on the current stream, and return their flat range for :meth:`all_reduce_mean` to skip.  ``None``
        (nothing done) when the parameters are not contiguous in the flat buffer."""
        index = {id(p): i for i, p in enumerate(self._plist)}
        idx = sorted(index[id(p)] for p in params if id(p) in index)
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            return None
        self._adopt_parameters()
        for i in idx:
            p, v = self._plist[i], self._gviews[i]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v
        start = self._offsets[idx[0]]
        stop = self._offsets[idx[-1]] + (self._plist[idx[-1]].numel() + 3) // 4 * 4
        self._all_reduce_avg(self.flat_grad[start:stop])
        return start, stop

    @torch.no_grad()
    def all_reduce_mean(self, skip: Optional[Tuple[int, int]] = None) -> None:
        r"""Data-parallel gradient exchange on the flat buffer (one message; two around ``skip``, a flat
        range that :meth:`all_reduce_subset` has exchanged already).  Ranks must agree on the message, so
        parameters without a gradient on this rank contribute zeros and take part in the update."""
        for v, p in zip(self._gviews, self._plist):
            if p.requires_grad and p.grad is None:
                v.zero_()
                p.grad = v
        self.pack_grads()
        if skip is None:
            self._all_reduce_avg(self.flat_grad)
        else:
            self._all_reduce_avg(self.flat_grad[:skip[0]])
            self._all_reduce_avg(self.flat_grad[skip[1]:])

    # ---------------------------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        ranges = self._packed if self._packed is not None else self.pack_grads()
        self._packed = None
        stream = current_stream(self.device)
        L = _lib.lib()
        base_p, base_g, base_s = ptr(self.flat_param), ptr(self.flat_grad), ptr(self.flat_sq)
        for start, stop in ranges:
            check(L.glue_rmsprop_flat(base_p + 4 * start, base_g + 4 * start, base_s + 4 * start,
                                      stop - start, float(group["lr"]), float(group["alpha"]),
                                      float(group["eps"]), stream), "rmsprop_flat")
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._packed = None
        super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._plist = self.param_groups[0]["params"]
        for p, sq in zip(self._plist, self._sviews):  # loaded tensors are copies: fold them back
            loaded = self.state[p].get("square_avg")
            if loaded is not None and loaded.data_ptr() != sq.data_ptr():
                sq.copy_(loaded)
            self.state[p]["square_avg"] = sq
        self._adopt_parameters()
