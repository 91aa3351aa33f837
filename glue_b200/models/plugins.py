r"""
Training plugins (API of ``scglue/models/plugins.py``): early stopping with best-checkpoint
restore, learning-rate reduction on plateau, optional Tensorboard logging.
"""

import pathlib
import re
from typing import List

import torch
from torch.optim.lr_scheduler import ReduceLROnPlateau

from .. import dist as gdist
from ..utils import config, logged
from .base import COMPLETED, EPOCH_COMPLETED, TERMINATE, Engine, TrainingPlugin


@logged
class Tensorboard(TrainingPlugin):
    r"""Per-epoch scalar logging (``plugins.py:23-60``).  A no-op when ``tensorboardX`` /
    ``torch.utils.tensorboard`` are unavailable."""

    def attach(self, net, trainer, train_engine, val_engine, train_loader, val_loader,
               directory) -> None:
        if gdist.rank() != 0:  # data parallel: metrics are rank-averaged, one writer is enough
            return
        try:
            from torch.utils.tensorboard import SummaryWriter
        except Exception:  # pragma: no cover
            self.logger.debug("tensorboard is not installed; scalar logging disabled")
            return
        log_dir = pathlib.Path(directory) / "tensorboard"
        writer = SummaryWriter(log_dir=str(log_dir), flush_secs=config.TENSORBOARD_FLUSH_SECS)

        def _log(engine, tag, source):
            for key, val in source.state.metrics.items():
                writer.add_scalar(f"{tag}/{key}", val, train_engine.state.epoch)

        train_engine.add_event_handler(EPOCH_COMPLETED, _log, "train", train_engine)
        if val_engine:
            train_engine.add_event_handler(EPOCH_COMPLETED, _log, "val", val_engine)
        train_engine.add_event_handler(COMPLETED, lambda engine: writer.close())
        train_engine.add_event_handler(TERMINATE, lambda engine: writer.close())


@logged
class EarlyStopping(TrainingPlugin):
    r"""
    Stop when the monitored loss has not improved for ``patience`` epochs, keeping the
    ``config.CHECKPOINT_SAVE_NUMBERS`` best checkpoints (``{"net", "trainer"}`` state dicts
    in ``checkpoint_{epoch}.pt``) and restoring the most recent of them when training ends;
    a checkpoint written in an epoch that ended with non-finite losses is discarded
    (``plugins.py:63-170``).  ``burnin`` epochs and ``wait_n_lrs`` learning-rate reductions
    are waited out first.
    """

    def __init__(self, monitor: str, patience: int, burnin: int = 0, wait_n_lrs: int = 0) -> None:
        super().__init__()
        self.monitor, self.patience = monitor, patience
        self.burnin, self.wait_n_lrs = burnin, wait_n_lrs

    def attach(self, net, trainer, train_engine, val_engine, train_loader, val_loader,
               directory) -> None:
        # Data parallel: parameters and optimizer state are identical on every rank (averaged
        # gradients) and so are the rank-averaged metrics, hence every rank takes the same keep /
        # stop decisions; rank 0 alone touches the files, all ranks restore the same checkpoint.
        directory = pathlib.Path(directory)
        writer = gdist.rank() == 0
        if writer:
            directory.mkdir(parents=True, exist_ok=True)
            for item in directory.glob("checkpoint_*.pt"):
                item.unlink()
        gdist.barrier()
        score_engine = val_engine if val_engine else train_engine
        saved: List[tuple] = []  # (score, epoch)
        tracker = {"best": None, "stale": 0}

        def _on_epoch(engine: Engine) -> None:
            epoch = engine.state.epoch
            if epoch <= self.burnin:
                return
            if self.wait_n_lrs and getattr(engine.state, "n_lrs", 0) < self.wait_n_lrs:
                return
            score = -score_engine.state.metrics[self.monitor]
            # keep the n best checkpoints (ties: the newer one wins, like ignite's Checkpoint)
            if len(saved) < config.CHECKPOINT_SAVE_NUMBERS or score >= min(saved)[0]:
                if writer:
                    torch.save({"net": net.state_dict(), "trainer": trainer.state_dict()},
                               directory / f"checkpoint_{epoch}.pt")
                saved.append((score, epoch))
                if len(saved) > config.CHECKPOINT_SAVE_NUMBERS:
                    worst = min(saved)
                    saved.remove(worst)
                    if writer:
                        (directory / f"checkpoint_{worst[1]}.pt").unlink(missing_ok=True)
            if tracker["best"] is None or score > tracker["best"]:
                tracker["best"], tracker["stale"] = score, 0
            else:
                tracker["stale"] += 1
                if tracker["stale"] >= self.patience:
                    self.logger.info("EarlyStopping: Stop training")
                    engine.terminate()

        def _on_end(engine: Engine) -> None:
            nan_flag = any(v != v or abs(v) == float("inf")
                           for v in engine.state.metrics.values())
            gdist.barrier()  # rank 0 has finished writing
            ckpts = sorted((int(re.match(r"checkpoint_(\d+)\.pt", p.name).group(1))
                            for p in directory.glob("checkpoint_*.pt")), reverse=True)
            if ckpts and nan_flag and engine.state.epoch == ckpts[0]:
                self.logger.warning('The most recent checkpoint "%d" can be corrupted by NaNs, '
                                    "will thus be discarded.", ckpts[0])
                ckpts = ckpts[1:]
            if ckpts:
                self.logger.info('Restoring checkpoint "%d"...', ckpts[0])
                loaded = torch.load(directory / f"checkpoint_{ckpts[0]}.pt",
                                    map_location=net.device, weights_only=False)
                net.load_state_dict(loaded["net"])
                trainer.load_state_dict(loaded["trainer"])
            else:
                self.logger.info("No usable checkpoint found. Skipping checkpoint restoration.")

        train_engine.add_event_handler(EPOCH_COMPLETED, _on_epoch)
        train_engine.add_event_handler(COMPLETED, _on_end)
        train_engine.add_event_handler(TERMINATE, _on_end)


@logged
class LRScheduler(TrainingPlugin):
    r"""``ReduceLROnPlateau`` on every optimizer, in lock-step (``plugins.py:173-243``)."""

    def __init__(self, *optims: torch.optim.Optimizer, monitor: str = None, patience: int = None,
                 burnin: int = 0) -> None:
        super().__init__()
        if monitor is None:
            raise ValueError("`monitor` must be specified!")
        if patience is None:
            raise ValueError("`patience` must be specified!")
        self.monitor = monitor
        self.schedulers = [ReduceLROnPlateau(optim, patience=patience) for optim in optims]
        self.burnin = burnin

    def attach(self, net, trainer, train_engine, val_engine, train_loader, val_loader,
               directory) -> None:
        score_engine = val_engine if val_engine else train_engine
        for scheduler in self.schedulers:
            scheduler.last_epoch = self.burnin
        train_engine.state.n_lrs = 0

        def _on_epoch(engine: Engine) -> None:
            if engine.state.epoch <= self.burnin:
                return
            changed = set()
            for scheduler in self.schedulers:
                old_lr = scheduler.optimizer.param_groups[0]["lr"]
                scheduler.step(score_engine.state.metrics[self.monitor])
                changed.add(scheduler.optimizer.param_groups[0]["lr"] != old_lr)
            if len(changed) != 1:
                raise RuntimeError("Learning rates are out of sync!")
            if changed.pop():
                engine.state.n_lrs += 1
                self.logger.info("Learning rate reduction: step %d", engine.state.n_lrs)

        train_engine.add_event_handler(EPOCH_COMPLETED, _on_epoch)
