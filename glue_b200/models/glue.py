r"""
GLUE network container and base trainer (API of ``scglue/models/glue.py``).
"""

import itertools
import os
from typing import Any, List, Mapping, Optional

import numpy as np
import torch

from .. import dist as gdist
from ..num import normalize_edges
from ..utils import config, logged
from .base import Trainer, TrainingPlugin
from .data import ArrayDataset, DataLoader, DevicePrefetcher, GraphDataset, ParallelDataLoader
from .nn import autodevice
from .plugins import EarlyStopping, LRScheduler, Tensorboard

# ------------------------------------------------------------- module interfaces


class GraphEncoder(torch.nn.Module):
    r"""``forward(eidx, enorm, esgn) -> distribution over vertex embeddings``."""


class GraphDecoder(torch.nn.Module):
    r"""``forward(v, eidx, esgn) -> distribution over edge existence``."""


class DataEncoder(torch.nn.Module):
    r"""``forward(x, ...) -> distribution over cell embeddings``."""


class DataDecoder(torch.nn.Module):
    r"""``forward(u, v, ...) -> distribution over observations``."""


class Discriminator(torch.nn.Module):
    r"""``forward(u, ...) -> modality logits``."""


class Prior(torch.nn.Module):
    r"""``forward() -> prior distribution``."""


# -------------------------------------------------------------------- container


class GLUE(torch.nn.Module):
    r"""
    Container of the GLUE sub-networks (``scglue/models/glue.py:177-243``): graph encoder
    ``g2v`` / decoder ``v2g``, per-modality data encoders ``x2u`` / decoders ``u2x``, the
    feature-to-vertex index buffers ``{k}_idx``, discriminator ``du`` and ``prior``.
    """

    def __init__(self, g2v, v2g, x2u: Mapping[str, torch.nn.Module],
                 u2x: Mapping[str, torch.nn.Module], idx: Mapping[str, torch.Tensor], du,
                 prior) -> None:
        super().__init__()
        if not set(x2u.keys()) == set(u2x.keys()) == set(idx.keys()) != set():
            raise ValueError("`x2u`, `u2x`, `idx` should share the same keys and non-empty!")
        self.keys = list(idx.keys())
        self.g2v, self.v2g = g2v, v2g
        self.x2u, self.u2x = torch.nn.ModuleDict(x2u), torch.nn.ModuleDict(u2x)
        for k, v in idx.items():
            self.register_buffer(f"{k}_idx", v)
        self.du, self.prior = du, prior
        self.device = autodevice()

    @property
    def device(self) -> torch.device:
        return self._device

    @device.setter
    def device(self, device: torch.device) -> None:
        self._device = device
        self.to(self._device)

    def forward(self):
        raise RuntimeError("GLUE does not support forward operation!")


# ---------------------------------------------------------------------- trainer


@logged
class GLUETrainer(Trainer):
    r"""
    Base trainer (``scglue/models/glue.py:257-740``): owns the two optimizers (``vae_optim``
    over g2v / v2g / x2u / u2x, ``dsc_optim`` over du), the loss weights and the data
    plumbing of :meth:`fit` / :meth:`get_losses`.

    Multi-GPU: when ``torch.distributed`` is initialised, every rank trains on its own
    cell minibatches and gradients are averaged with one flat-bucket NCCL all-reduce per
    optimizer step (see :mod:`glue_b200.dist`); the guidance graph and all parameters are
    replicated.
    """

    def __init__(self, net: GLUE, lam_data: float = None, lam_kl: float = None,
                 lam_graph: float = None, lam_align: float = None, dsc_steps: int = None,
                 modality_weight: Mapping[str, float] = None, optim: str = None,
                 lr: float = None, **kwargs) -> None:
        given = dict(lam_data=lam_data, lam_kl=lam_kl, lam_graph=lam_graph, lam_align=lam_align,
                     dsc_steps=dsc_steps, modality_weight=modality_weight, optim=optim, lr=lr)
        for name, val in given.items():
            if val is None:
                raise ValueError(f"`{name}` must be specified!")
        super().__init__(net)
        self.required_losses = ["g_nll", "g_kl", "g_elbo"]
        for k in net.keys:
            self.required_losses += [f"x_{k}_nll", f"x_{k}_kl", f"x_{k}_elbo"]
        self.required_losses += ["dsc_loss", "vae_loss", "gen_loss"]
        self.earlystop_loss = "vae_loss"

        self.lam_data, self.lam_kl, self.lam_graph = lam_data, lam_kl, lam_graph
        self.lam_align, self.dsc_steps = lam_align, dsc_steps
        if min(modality_weight.values()) < 0:
            raise ValueError("Modality weight must be non-negative!")
        normalizer = sum(modality_weight.values()) / len(modality_weight)
        self.modality_weight = {k: v / normalizer for k, v in modality_weight.items()}

        self.lr = lr
        self.optim_name, self.optim_kwargs = optim, dict(kwargs)
        self.vae_optim = self._make_optim(self._vae_modules())
        self.dsc_optim = self._make_optim([net.du])
        self.align_burnin: Optional[int] = None
        self.eidx = self.enorm = self.esgn = None  # full graph used by the graph encoder
        self.full_graph_digest: Optional[str] = None
        self._vae_bucket = self._dsc_bucket = None

    def _vae_modules(self) -> List[torch.nn.Module]:
        net = self.net
        return [net.g2v, net.v2g, net.x2u, net.u2x]

    def _make_optim(self, modules) -> torch.optim.Optimizer:
        params = list(itertools.chain(*(m.parameters() for m in modules)))
        kwargs = dict(self.optim_kwargs)
        on_gpu = self.net.device.type == "cuda"
        if (on_gpu and self.optim_name == "RMSprop" and config.USE_FUSED_OPTIMIZER
                and set(kwargs) <= {"alpha", "eps"}):
            # the reference's optimizer (scglue.py:889-942) as one fused kernel per step
            from ..optim import FlatRMSprop
            return FlatRMSprop(params, lr=self.lr, **kwargs)
        if on_gpu and self.optim_name in ("RMSprop", "Adam", "AdamW"):
            kwargs.setdefault("capturable", True)  # same arithmetic, no host reads: graph-safe
        return getattr(torch.optim, self.optim_name)(params, lr=self.lr, **kwargs)

    # -- data-parallel gradient exchange -------------------------------------------------
    def _sync_grads(self, which: str, skip=None) -> None:
        r"""``skip``: flat range of the optimizer's gradient buffer that has been exchanged already
        (``FlatRMSprop.all_reduce_subset``)."""
        if not gdist.is_distributed():
            return
        optim = self.vae_optim if which == "vae" else self.dsc_optim
        if hasattr(optim, "all_reduce_mean"):  # FlatRMSprop: gradients already live in one buffer
            optim.all_reduce_mean(skip=skip)
            return
        attr = f"_{which}_bucket"
        if getattr(self, attr) is None:
            optim = self.vae_optim if which == "vae" else self.dsc_optim
            params = [p for g in optim.param_groups for p in g["params"] if p.requires_grad]
            setattr(self, attr, gdist.GradBucket(params))
        getattr(self, attr).all_reduce_mean()

    def set_full_graph(self, graph: GraphDataset) -> None:
        r"""Normalise the full guidance graph and keep it on the device
        (``glue.py:555-559``); cleared again at the end of :meth:`fit`.

        The device tensors of the most recent graph are kept behind the scenes, keyed by a digest
        of its edges: captured training steps have the CSR arrays of a specific graph baked in
        (``full_graph_digest`` is part of their key), and a second ``fit`` on the same graph finds
        the same tensors -- same CSR cache entry, same captures -- instead of re-uploading."""
        import hashlib
        device = self.net.device
        enorm = normalize_edges(graph.eidx, graph.ewt)
        h = hashlib.blake2b(digest_size=16)
        for arr in (graph.eidx, enorm, graph.esgn):
            h.update(np.ascontiguousarray(arr).tobytes())
        digest = h.hexdigest()
        kept = getattr(self, "_full_graph_kept", None)
        if kept is None or kept[0] != (digest, str(device)):
            kept = ((digest, str(device)), torch.as_tensor(graph.eidx, device=device),
                    torch.as_tensor(enorm, device=device), torch.as_tensor(graph.esgn, device=device))
            self._full_graph_kept = kept
        (self.full_graph_digest, _), self.eidx, self.enorm, self.esgn = kept

    def clear_full_graph(self) -> None:
        self.eidx = self.enorm = self.esgn = None
        self.full_graph_digest = None

    # -- fit / get_losses ------------------------------------------------------------------
    def _loaders(self, data, graph, data_batch: int, graph_batch: int, random_seed: int,
                 drop_last_data: bool, drop_last_graph: bool, shard: bool = False):
        r"""Block-shuffled data loader zipped with a cycling graph loader
        (``glue.py:580-625``).  Samplers are built explicitly so that the generator
        consumption equals ``DataLoader(shuffle=True, generator=G)`` and so that, under data
        parallelism (``shard``), rank ``r`` materialises only batches ``r, r + N, ...`` of
        the common batch sequence while every rank sees the same edge minibatches."""
        def make(ds, batch, drop_last, strided):
            on_device = getattr(ds, "device", None) is not None
            if hasattr(ds, "sparse_items"):  # host-resident count matrices feeding a GPU travel sparse
                ds.sparse_items = (config.SPARSE_HOST_BATCHES and not on_device
                                   and self.net.device.type == "cuda")
            pin = (config.DATALOADER_PIN_MEMORY and not config.CPU_ONLY and not on_device
                   and torch.cuda.is_available())
            gen = torch.Generator().manual_seed(random_seed)
            batches = torch.utils.data.BatchSampler(
                torch.utils.data.RandomSampler(ds, generator=gen), batch, drop_last)
            if strided and gdist.is_distributed():
                batches = gdist.RankStridedBatches(batches, gdist.rank(), gdist.world_size())
            return DataLoader(ds, reshuffle=True, batch_sampler=batches, generator=gen,
                              num_workers=config.DATALOADER_NUM_WORKERS, pin_memory=pin,
                              persistent_workers=False)
        return ParallelDataLoader(make(data, data_batch, drop_last_data, shard),
                                  make(graph, graph_batch, drop_last_graph, False),
                                  cycle_flags=[False, True])

    def fit(self, data: ArrayDataset, graph: GraphDataset, val_split: float = None,
            data_batch_size: int = None, graph_batch_size: int = None, align_burnin: int = None,
            safe_burnin: bool = True, max_epochs: int = None, patience: Optional[int] = None,
            reduce_lr_patience: Optional[int] = None, wait_n_lrs: Optional[int] = None,
            random_seed: int = None, directory: Optional[os.PathLike] = None,
            plugins: Optional[List[TrainingPlugin]] = None) -> None:
        r"""
        Fit the network (``glue.py:486-667``): 90/10 split of the view rows, block-shuffled
        loaders (4 fetches of ``batch/4`` rows per minibatch), the graph loader cycles,
        learning-rate scheduling and early stopping on the validation ``vae_loss``.
        """
        given = dict(val_split=val_split, data_batch_size=data_batch_size,
                     graph_batch_size=graph_batch_size, align_burnin=align_burnin,
                     max_epochs=max_epochs, random_seed=random_seed)
        for name, val in given.items():
            if val is None:
                raise ValueError(f"`{name}` must be specified!")
        if patience and reduce_lr_patience and reduce_lr_patience >= patience:
            self.logger.warning("Parameter `reduce_lr_patience` should be smaller than "
                                "`patience`, otherwise learning rate scheduling would be "
                                "ineffective.")
        self.set_full_graph(graph)
        fetches = config.DATALOADER_FETCHES_PER_BATCH
        data.getitem_size = max(1, round(data_batch_size / fetches))
        graph.getitem_size = max(1, round(graph_batch_size / fetches))
        data_train, data_val = data.random_split([1 - val_split, val_split],
                                                 random_state=random_seed)
        data_train.prepare_shuffle(num_workers=config.ARRAY_SHUFFLE_NUM_WORKERS,
                                   random_seed=random_seed)
        data_val.prepare_shuffle(num_workers=config.ARRAY_SHUFFLE_NUM_WORKERS,
                                 random_seed=random_seed)
        graph.prepare_shuffle(num_workers=config.GRAPH_SHUFFLE_NUM_WORKERS,
                              random_seed=random_seed)
        train_loader = self._loaders(data_train, graph, fetches, fetches, random_seed,
                                     len(data_train) > fetches, len(graph) > fetches, shard=True)
        val_loader = self._loaders(data_val, graph, fetches, fetches, random_seed, False, False)
        if self.net.device.type == "cuda" and getattr(data, "device", None) is None:
            # host-resident datasets: upload minibatch i + 1 while step i runs
            train_loader = DevicePrefetcher(train_loader, self.net.device)
        self.align_burnin = align_burnin

        default_plugins: List[TrainingPlugin] = [Tensorboard()]
        burnin = self.align_burnin if safe_burnin else 0
        if reduce_lr_patience:
            default_plugins.append(LRScheduler(self.vae_optim, self.dsc_optim,
                                               monitor=self.earlystop_loss,
                                               patience=reduce_lr_patience, burnin=burnin))
        if patience:
            default_plugins.append(EarlyStopping(monitor=self.earlystop_loss, patience=patience,
                                                 burnin=burnin, wait_n_lrs=wait_n_lrs or 0))
        try:
            super().fit(train_loader, val_loader=val_loader, max_epochs=max_epochs,
                        random_seed=random_seed, directory=directory,
                        plugins=default_plugins + (plugins or []))
        finally:
            for ds in (data, data_train, data_val, graph):
                ds.clean()
            self.align_burnin = None
            self.clear_full_graph()

    def get_losses(self, data: ArrayDataset, graph: GraphDataset, data_batch_size: int = None,
                   graph_batch_size: int = None, random_seed: int = None) -> Mapping[str, float]:
        for name, val in dict(data_batch_size=data_batch_size, graph_batch_size=graph_batch_size,
                              random_seed=random_seed).items():
            if val is None:
                raise ValueError(f"`{name}` must be specified!")
        self.set_full_graph(graph)
        data.getitem_size, graph.getitem_size = data_batch_size, graph_batch_size
        data.prepare_shuffle(num_workers=config.ARRAY_SHUFFLE_NUM_WORKERS, random_seed=random_seed)
        graph.prepare_shuffle(num_workers=config.GRAPH_SHUFFLE_NUM_WORKERS,
                              random_seed=random_seed)
        loader = self._loaders(data, graph, 1, 1, random_seed, False, False)
        try:
            return super().get_losses(loader)
        finally:
            data.clean()
            graph.clean()
            self.clear_full_graph()

    def state_dict(self) -> Mapping[str, Any]:
        return {**super().state_dict(), "vae_optim": self.vae_optim.state_dict(),
                "dsc_optim": self.dsc_optim.state_dict()}

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        state_dict = dict(state_dict)
        self.vae_optim.load_state_dict(state_dict.pop("vae_optim"))
        self.dsc_optim.load_state_dict(state_dict.pop("dsc_optim"))
        super().load_state_dict(state_dict)
        graphs = getattr(self, "graphs", None)
        if graphs is not None:  # captured steps reference the replaced optimizer state
            graphs.invalidate()
