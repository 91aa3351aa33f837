r"""
CUDA-graph capture / replay of whole training steps.

One SCGLUE ``train_step`` is ~700 small launches (two forward/backward passes, two RMSprop
updates) around a handful of fused kernels; eagerly it is bound by the host launch path, not
by the GPU.  After a couple of eager steps with a given input signature the step is captured
once -- discriminator update, generator update, both optimizers, the gradient all-reduce when
data parallel -- and afterwards every step is: copy the minibatch into the static input
buffers, one graph launch.  Nothing about the arithmetic changes: the captured kernels are the
ones the eager step launches, in the same order, on the same parameters and optimizer state.

What makes the step capturable (no host synchronisation, static shapes):

* the graph-reconstruction loss balances positive / negative edges on the device
  (``ops.graph_nll``; the reference reads the positive count with ``.item()``,
  ``scglue.py:333``);
* the paired-cell losses use per-cell weights instead of compacted row subsets
  (``PairedSCGLUETrainer._extra_losses``);
* RMSprop runs with ``capturable=True``; the learning rate is part of the graph key, so a
  ``ReduceLROnPlateau`` step simply leads to a re-capture;
* the burn-in noise coefficient is a device scalar refreshed before every replay.
"""

import gc
from collections import OrderedDict
from typing import Any, Dict, List, Mapping, Optional

import torch

from .. import ops
from ..utils import config, logged
from .base import LossVector


@logged
class StepGraphs:
    r"""Per-trainer cache of captured training steps, keyed by everything that is baked into a
    capture: input shapes / dtypes, the ``freeze_u`` / burn-in switches and the learning rates."""

    WARMUP_STEPS = 2        # eager steps before the first capture (optimizer state, workspaces)
    WARMUP_STEPS_LATER = 1  # eager steps with a further key (e.g. the short last minibatch)
    MAX_GRAPHS = 8          # training / validation steps x (full, short last) minibatch x burn-in on / off

    def __init__(self, trainer) -> None:
        self.trainer = trainer
        self.entries: "OrderedDict[tuple, Dict[str, Any]]" = OrderedDict()
        self.pool = None
        self.disabled_reason: Optional[str] = None
        self.replays = 0

    # ------------------------------------------------------------------------------ helpers
    def usable(self) -> bool:
        tr = self.trainer
        return (self.disabled_reason is None and config.USE_CUDA_GRAPHS
                and tr.net.device.type == "cuda" and not ops.TIMER.enabled)

    def invalidate(self) -> None:
        r"""Drop every capture (optimizer state or parameters were replaced, e.g. by
        ``load_state_dict``)."""
        self.entries.clear()
        self.pool = None  # the memory pool dies with its last graph: later captures get a new one

    def _key(self, data: List[torch.Tensor], anneal: float) -> tuple:
        tr = self.trainer
        lrs = tuple(float(g["lr"]) for opt in (tr.vae_optim, tr.dsc_optim) for g in opt.param_groups)
        # loss weights are Python numbers baked into the captured kernels' arguments
        weights = tuple(float(getattr(tr, name)) for name in (
            "lam_data", "lam_kl", "lam_graph", "lam_align", "lam_sup", "lam_joint_cross",
            "lam_real_cross", "lam_cos") if hasattr(tr, name))
        weights += tuple(float(v) for v in tr.modality_weight.values())
        # (the guidance graph's CSR arrays are baked into a capture: another graph is another key)
        return (tuple((tuple(t.shape), t.dtype) for t in data), bool(anneal), bool(tr.freeze_u),
                lrs, tr.dsc_steps, weights, bool(tr.normalize_u), tr.align_burnin,
                getattr(tr, "full_graph_digest", None))

    def _storage_fingerprint(self) -> tuple:
        r"""Addresses of the first and last parameter of every optimizer: a capture is only valid
        while the parameters live where they lived when it was taken (``net.to(...)``, as done by
        ``Model.save``, re-allocates them)."""
        tr = self.trainer
        out = []
        for opt in (tr.vae_optim, tr.dsc_optim):
            params = opt.param_groups[0]["params"]
            out += [params[0].data_ptr(), params[-1].data_ptr()]
        return tuple(out)

    def info(self) -> Mapping[str, Any]:
        r"""``{"graphs", "replays", "kernels_per_step"}`` (native kernels inside the most
        recently used capture)."""
        last = next(reversed(self.entries.values()), None) if self.entries else None
        return {"graphs": sum(1 for e in self.entries.values() if e["graph"] is not None),
                "replays": self.replays,
                "kernels_per_step": None if last is None else last.get("native_kernels"),
                "disabled": self.disabled_reason}

    # ------------------------------------------------------------------------------ stepping
    def step(self, engine, data: List[torch.Tensor], mode: str = "train") -> Mapping[str, torch.Tensor]:
        r"""``mode = "train"``: one training step; ``"val"``: one validation step (eval-mode forward pass
        under ``no_grad``; eagerly that is ~300 launches bound by the host launch path, 10 ms for a step the
        device finishes in 0.4 ms -- two thirds of an epoch of ``fit`` at the PBMC shape)."""
        tr = self.trainer
        eager = tr.eager_train_step if mode == "train" else tr.eager_val_step
        fingerprint = self._storage_fingerprint()
        if fingerprint != getattr(self, "_fingerprint", fingerprint):
            self.logger.info("parameters were re-allocated: dropping captured training steps")
            self.invalidate()
        self._fingerprint = fingerprint
        anneal = tr.burnin_anneal(engine.state.epoch)
        key = self._key(data, anneal) + (mode,)
        entry = self.entries.get(key)
        if entry is None:
            entry = {"seen": 0, "graph": None}
            self.entries[key] = entry
            while len(self.entries) > self.MAX_GRAPHS:
                self.entries.popitem(last=False)
        else:
            self.entries.move_to_end(key)
        if entry["graph"] is None:
            entry["seen"] += 1
            have_graph = any(e["graph"] is not None for e in self.entries.values())
            if entry["seen"] <= (self.WARMUP_STEPS_LATER if have_graph else self.WARMUP_STEPS):
                # warm-up steps take the code path of the capture (device-scalar burn-in
                # coefficient): every kernel the capture launches has been loaded by then -- a lazy
                # module load inside a capture invalidates it
                tr.anneal_buffer.fill_(anneal)
                out = eager(engine, data)
                self._fingerprint = self._storage_fingerprint()  # the optimizer may re-adopt storage
                return out
            tr.anneal_buffer.fill_(anneal)
            self._capture(entry, engine, data, eager)
            if entry["graph"] is None:
                return eager(engine, data)
        self._load_inputs(entry["inputs"], data)
        tr.anneal_buffer.fill_(anneal)
        entry["graph"].replay()
        self.replays += 1
        return entry["outputs"]

    @staticmethod
    def _load_inputs(statics: List[torch.Tensor], data: List[torch.Tensor]) -> None:
        r"""Minibatch -> static input buffers of the captured step.  Tensors that already live on the
        device go through ONE launch (``ops.multi_copy``; a paired PBMC step has 14 inputs: as
        individual copies they were 60 us of launch latency in front of every replay); host tensors
        are copied one by one (asynchronously when pinned)."""
        dsts, srcs = [], []
        for static, fresh in zip(statics, data):
            if (fresh.device == static.device and fresh.dtype == static.dtype and fresh.shape == static.shape
                    and fresh.is_contiguous() and static.is_contiguous() and static.numel()):
                dsts.append(static), srcs.append(fresh)
            else:
                static.copy_(fresh, non_blocking=True)
        if len(dsts) > 1:
            ops.multi_copy(dsts, srcs)
        elif dsts:
            dsts[0].copy_(srcs[0], non_blocking=True)

    @staticmethod
    def _recover_generator(device: torch.device, state: Optional[torch.Tensor]) -> None:
        r"""An aborted capture leaves the device's default generator with a non-zero intra-graph
        offset, and every later random draw raises.  Hand it the state object of a fresh generator
        (same seed and offset as before the capture) so that eager training can go on."""
        if state is None:
            return
        try:
            fresh = torch.Generator(device=device)
            fresh.set_state(state)
            index = device.index if device.index is not None else torch.cuda.current_device()
            torch.cuda.default_generators[index].graphsafe_set_state(fresh.graphsafe_get_state())
        except Exception:  # best effort: the original error is what gets reported
            pass

    def _capture(self, entry: Dict[str, Any], engine, data: List[torch.Tensor], eager) -> None:
        tr = self.trainer
        device = tr.net.device
        try:
            inputs = [torch.empty(t.shape, dtype=t.dtype, device=device) for t in data]
            for static, fresh in zip(inputs, data):
                static.copy_(fresh, non_blocking=True)
            torch.cuda.synchronize(device)
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            # whatever the step caches across calls must exist before the capture: memory allocated
            # inside it belongs to the graph's pool and is only meaningful during a replay
            tr.prepare_capture()
            rng_before = torch.cuda.get_rng_state(device)  # (to recover the generator if the capture fails)
            graph = torch.cuda.CUDAGraph()
            launches0 = ops.kernel_launches()
            tr.capturing = True
            # no cyclic garbage collection while the capture is open: a finalizer that releases CUDA memory or
            # an older graph (a trainer of an earlier fit waiting in a reference cycle) is an operation the
            # capturing thread may not perform, and it would invalidate the capture
            gc_was_on = gc.isenabled()
            gc.collect()
            gc.disable()
            try:
                with torch.cuda.graph(graph, pool=self.pool, capture_error_mode="thread_local"):
                    outputs = eager(engine, inputs)
            finally:
                tr.capturing = False
                if gc_was_on:
                    gc.enable()
            entry.update(graph=graph, inputs=inputs,
                         outputs=outputs if isinstance(outputs, LossVector) else dict(outputs),
                         native_kernels=ops.kernel_launches() - launches0)
            self.logger.debug("captured a training step (%d native kernels)",
                              entry["native_kernels"])
        except Exception as e:  # capture is an optimisation: fall back to eager launches
            torch.cuda.synchronize(device)
            self._recover_generator(device, locals().get("rng_before"))
            import traceback
            self.disabled_reason = f"{type(e).__name__}: {e}"
            self.logger.debug("capture traceback:\n%s", traceback.format_exc())
            self.last_traceback = traceback.format_exc()
            self.logger.warning("CUDA-graph capture of the training step failed (%s); "
                                "continuing with eager launches\n%s", self.disabled_reason,
                                "".join(self.last_traceback.splitlines(keepends=True)[-24:]))
            entry["graph"] = None
