r"""
Training loop, model persistence and the plugin interface
(API of ``scglue/models/base.py``, without the pytorch-ignite dependency).

The reference wires two ``ignite.engine.Engine`` objects (``base.py:117-214``); here a
small event-driven :class:`Engine` provides the same hooks (``state.epoch``,
``state.metrics``, ``terminate()``, per-epoch validation, plugins).  Two deliberate
differences, both to keep the GPU queue free of host synchronisation:

* running loss averages are accumulated **on the device** and read back once per epoch
  (ignite's ``Average`` reads every step);
* the non-finite-loss check (``TerminateOnNan``) therefore fires at the end of the
  offending epoch instead of at the offending iteration;
* the read-back itself waits until somebody looks at the numbers (:class:`DeferredMetrics`), and the
  validation pass is issued before anybody does: its steps queue up behind the last training steps
  instead of starting on an idle device.
"""

import os
import pathlib
import tempfile
import time
from collections import defaultdict
from typing import Any, Callable, Dict, Iterable, List, Mapping, Optional

import dill
import torch

from ..utils import config, logged

# ----------------------------------------------------------------------- engine

STARTED, EPOCH_STARTED, ITERATION_COMPLETED = "started", "epoch_started", "iteration_completed"
EPOCH_COMPLETED, COMPLETED, TERMINATE = "epoch_completed", "completed", "terminate"


class EngineState:
    def __init__(self) -> None:
        self.epoch = 0
        self.iteration = 0
        self.max_epochs = 0
        self.output: Optional[Mapping[str, torch.Tensor]] = None
        self.metrics: Dict[str, float] = {}
        self.times: Dict[str, float] = {}


class LossVector(Mapping):
    r"""``{name: 0-dim tensor}`` backed by one device vector: the loss terms of a step come out of
    one matrix-vector product (``SCGLUETrainer._combine``), and the engine averages them over an
    epoch with one vector add per step instead of one add per term."""

    def __init__(self, names: List[str], values: torch.Tensor) -> None:
        self.names, self.values = list(names), values
        self._index = {name: i for i, name in enumerate(self.names)}

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.values[self._index[name]]

    def __iter__(self):
        return iter(self.names)

    def __len__(self) -> int:
        return len(self.names)

    def index(self, name: str) -> int:
        return self._index[name]

    def detach(self) -> "LossVector":
        return LossVector(self.names, self.values.detach())


class DeferredMetrics(Mapping):
    r"""Epoch means ``{name: float}`` that stay on the device until the first read (one device -> host copy,
    one synchronisation, then a plain dict)."""

    def __init__(self, tracked: List[str], names: List[str], means: torch.Tensor) -> None:
        self._tracked, self._names, self._means, self._host = list(tracked), list(names), means, None

    def resolve(self) -> Dict[str, float]:
        if self._host is None:
            host = self._means.tolist()
            self._host = {key: host[self._names.index(key)] for key in self._tracked}
            self._means = None
        return self._host

    def __getitem__(self, name: str) -> float:
        return self.resolve()[name]

    def __iter__(self):
        return iter(self._tracked)

    def __len__(self) -> int:
        return len(self._tracked)

    def __repr__(self) -> str:
        return repr(self.resolve())


class Engine:
    r"""Runs ``step(engine, batch)`` over a loader for a number of epochs and fires events."""

    def __init__(self, step: Callable[["Engine", Any], Mapping[str, torch.Tensor]],
                 tracked: Optional[List[str]] = None, sync_metrics: bool = False) -> None:
        self.step = step
        self.tracked = list(tracked or [])
        # data parallel: epoch metrics are averaged over the ranks so that every rank's plugins
        # (LR scheduling, early stopping, the non-finite guard) take the same decisions
        self.sync_metrics = sync_metrics
        self.state = EngineState()
        self.should_terminate = False
        self._handlers: Dict[str, List[Callable]] = defaultdict(list)

    def add_event_handler(self, event: str, handler: Callable, *args, **kwargs) -> None:
        self._handlers[event].append(lambda engine: handler(engine, *args, **kwargs))

    def on(self, event: str):
        def deco(fn):
            self.add_event_handler(event, fn)
            return fn
        return deco

    def fire(self, event: str) -> None:
        for handler in self._handlers[event]:
            handler(self)

    def terminate(self) -> None:
        self.should_terminate = True

    def run(self, loader: Iterable, max_epochs: int = 1) -> EngineState:
        r"""Run (or resume: epochs continue from ``state.epoch``) until ``max_epochs``."""
        state = self.state
        state.max_epochs = max_epochs
        if state.epoch == 0:
            self.fire(STARTED)
        while state.epoch < max_epochs and not self.should_terminate:
            state.epoch += 1
            tic = time.time()
            self.fire(EPOCH_STARTED)
            sums: Dict[str, torch.Tensor] = {}
            vec_sum, vec_names = None, None  # steps that return a LossVector: one add per step
            count = 0
            for batch in loader:
                state.iteration += 1
                state.output = self.step(self, batch)
                out = state.output
                if isinstance(out, LossVector) and self.tracked and (vec_names is None or vec_names == out.names) \
                        and not sums and all(key in out for key in self.tracked):
                    vec = out.values.detach()
                    vec_sum = vec.clone() if vec_sum is None else vec_sum + vec
                    vec_names = out.names
                else:
                    if vec_sum is not None:  # (mixed output kinds within an epoch: fold the vector in)
                        sums = {key: vec_sum[vec_names.index(key)].clone() for key in self.tracked}
                        vec_sum = None
                    for key in self.tracked:
                        val = out[key].detach()
                        sums[key] = val.clone() if key not in sums else sums[key] + val
                count += 1
                self.fire(ITERATION_COMPLETED)
                if self.should_terminate:
                    break
            if vec_sum is not None:  # one device->host read per epoch, when the numbers are looked at
                state.metrics = DeferredMetrics(self.tracked, vec_names, vec_sum.float() / max(count, 1))
            elif sums:
                keys = list(sums)
                host = (torch.stack([sums[k].float() for k in keys]) / max(count, 1)).tolist()
                state.metrics = dict(zip(keys, host))
            if self.sync_metrics and state.metrics:
                from .. import dist as gdist
                keys = list(state.metrics)
                state.metrics = dict(zip(keys, gdist.mean_over_ranks([state.metrics[k] for k in keys])))
            state.times["EPOCH_COMPLETED"] = time.time() - tic
            self.fire(EPOCH_COMPLETED)
            state.times["EPOCH_COMPLETED"] = time.time() - tic
        self.fire(TERMINATE if self.should_terminate else COMPLETED)
        return state


# ---------------------------------------------------------------------- trainer


@logged
class Trainer:
    r"""Abstract trainer: ``train_step`` / ``val_step`` + :meth:`fit` / :meth:`get_losses`."""

    def __init__(self, net: torch.nn.Module) -> None:
        self.net = net
        self.required_losses: List[str] = []

    def train_step(self, engine: Engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        raise NotImplementedError  # pragma: no cover

    def val_step(self, engine: Engine, data: List[torch.Tensor]) -> Mapping[str, torch.Tensor]:
        raise NotImplementedError  # pragma: no cover

    def report_metrics(self, train_state: EngineState, val_state: Optional[EngineState]) -> None:
        if train_state.epoch % config.PRINT_LOSS_INTERVAL:
            return
        fmt = lambda m: {k: float(f"{v:.3f}") for k, v in m.items()}
        self.logger.info("[Epoch %d] train=%s, val=%s, %.1fs elapsed", train_state.epoch,
                         fmt(train_state.metrics), fmt(val_state.metrics) if val_state else None,
                         train_state.times["EPOCH_COMPLETED"])

    def fit(self, train_loader: Iterable, val_loader: Optional[Iterable] = None,
            max_epochs: int = 100, random_seed: int = 0,
            directory: Optional[os.PathLike] = None,
            plugins: Optional[List["TrainingPlugin"]] = None) -> None:
        r"""Training loop with per-epoch validation and plugins (``base.py:117-214``)."""
        from .. import dist as gdist
        if directory is None:  # (data parallel: one directory, chosen by rank 0)
            directory = gdist.broadcast_object(
                tempfile.mkdtemp(prefix=config.TMP_PREFIX) if gdist.rank() == 0 else None)
        directory = pathlib.Path(directory)
        self.logger.info('Using training directory: "%s"', directory)
        sync = gdist.is_distributed()
        train_engine = Engine(self.train_step, self.required_losses, sync_metrics=sync)
        val_engine = Engine(self.val_step, self.required_losses, sync_metrics=sync) if val_loader else None

        if val_engine:  # (first: nothing has read the epoch's metrics yet, the device is still busy)
            @train_engine.on(EPOCH_COMPLETED)
            def _validate(engine):
                val_engine.run(val_loader, max_epochs=engine.state.epoch)

        @train_engine.on(EPOCH_COMPLETED)
        def _nonfinite_guard(engine):  # TerminateOnNan, at epoch granularity
            if any(v != v or v in (float("inf"), float("-inf"))
                   for v in engine.state.metrics.values()):
                self.logger.warning("Non-finite loss encountered, terminating training...")
                engine.terminate()

        @train_engine.on(EPOCH_COMPLETED)
        def _report(engine):
            self.report_metrics(engine.state, val_engine.state if val_engine else None)

        for plugin in plugins or []:
            plugin.attach(net=self.net, trainer=self, train_engine=train_engine,
                          val_engine=val_engine, train_loader=train_loader,
                          val_loader=val_loader, directory=directory)

        torch.manual_seed(random_seed)
        try:
            train_engine.run(train_loader, max_epochs=max_epochs)
        except KeyboardInterrupt:
            if not config.ALLOW_TRAINING_INTERRUPTION:
                raise
            self.logger.info("Stopping training due to user interrupt...")
            train_engine.terminate()
            train_engine.fire(TERMINATE)
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

    def get_losses(self, loader: Iterable) -> Mapping[str, float]:
        engine = Engine(self.val_step, self.required_losses)
        engine.run(loader, max_epochs=1)
        return dict(engine.state.metrics)

    def state_dict(self) -> Mapping[str, Any]:
        return {}

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        pass


# ------------------------------------------------------------------------ model


@logged
class Model:
    r"""Abstract model: owns a network, builds a trainer in :meth:`compile`, trains in
    :meth:`fit`, persists with ``dill`` (``base.py:261-388``)."""

    NET_TYPE = torch.nn.Module
    TRAINER_TYPE = Trainer

    def __init__(self, *args, **kwargs) -> None:
        self._net = self.NET_TYPE(*args, **kwargs)
        self._trainer: Optional[Trainer] = None

    @property
    def net(self) -> torch.nn.Module:
        return self._net

    @property
    def trainer(self) -> Trainer:
        if self._trainer is None:
            raise RuntimeError("No trainer has been registered! Please call `.compile()` first.")
        return self._trainer

    def compile(self, *args, **kwargs) -> None:
        if self._trainer:
            self.logger.warning("`compile` has already been called. "
                                "Previous trainer will be overwritten!")
        self._trainer = self.TRAINER_TYPE(self.net, *args, **kwargs)

    def fit(self, *args, **kwargs) -> None:
        self.trainer.fit(*args, **kwargs)

    def get_losses(self, *args, **kwargs) -> Mapping[str, float]:
        return self.trainer.get_losses(*args, **kwargs)

    def save(self, fname: os.PathLike) -> None:
        r"""``dill`` the model with the trainer stripped and the network moved to CPU."""
        fname = pathlib.Path(fname)
        trainer_backup, device_backup = self._trainer, self.net.device
        self._trainer, self.net.device = None, torch.device("cpu")
        try:
            with fname.open("wb") as f:
                dill.dump(self, f, protocol=4, byref=False, recurse=True)
        finally:
            self.net.device = device_backup
            self._trainer = trainer_backup
            graphs = getattr(trainer_backup, "graphs", None)
            if graphs is not None:  # the round trip through the CPU re-allocated every parameter
                graphs.invalidate()

    @staticmethod
    def load(fname: os.PathLike) -> "Model":
        fname = pathlib.Path(fname)
        with fname.open("rb") as f:
            model = dill.load(f)
        from .nn import autodevice
        model.net.device = autodevice()
        return model


@logged
class TrainingPlugin:
    r"""Plugin interface (``base.py:391-428``): ``attach`` receives the engines and loaders."""

    def attach(self, net: torch.nn.Module, trainer: Trainer, train_engine: Engine,
               val_engine: Optional[Engine], train_loader: Iterable,
               val_loader: Optional[Iterable], directory: pathlib.Path) -> None:
        raise NotImplementedError  # pragma: no cover
