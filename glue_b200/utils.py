r"""
Configuration, logging and RNG helpers (the subset of ``scglue/utils.py`` that the
training hot path reads: ``config``, ``log``, ``get_rs``, ``AUTO``).
"""

import logging
import os
import signal
from typing import Any, Optional, Union

import numpy as np
import torch

AUTO = "AUTO"  # sentinel: derive the hyper-parameter from the data (scglue/utils.py:21)

RandomState = Optional[Union[np.random.RandomState, int]]


class _Settings:
    r"""
    Global settings, same key names as the reference's ``ConfigManager``
    (``scglue/utils.py:180-451``) for the keys the hot path consults, plus the
    B200-specific knobs at the bottom.
    """

    _DEFAULTS = dict(
        TMP_PREFIX="GLUETMP",
        ANNDATA_KEY="__scglue__",
        CPU_ONLY=False,
        CUDNN_MODE="repeatability",
        MASKED_GPUS=[],
        ARRAY_SHUFFLE_NUM_WORKERS=0,
        # look-ahead threads of the negative sampler (reference default: 1 worker process).  Results
        # do not depend on it; a graph epoch is ~35 ms of GPU time at the PBMC shape and 0.18 s of
        # sampling, so a B200 needs several shuffles in flight
        GRAPH_SHUFFLE_NUM_WORKERS=4,
        FORCE_TERMINATE_WORKER_PATIENCE=60,
        DATALOADER_NUM_WORKERS=0,
        DATALOADER_FETCHES_PER_WORKER=4,
        DATALOADER_PIN_MEMORY=True,
        CHECKPOINT_SAVE_INTERVAL=10,
        CHECKPOINT_SAVE_NUMBERS=3,
        PRINT_LOSS_INTERVAL=10,
        TENSORBOARD_FLUSH_SECS=5,
        ALLOW_TRAINING_INTERRUPTION=True,
        # --- glue_b200 additions ---
        # keep count matrices resident in HBM (dense) when they fit this many bytes per GPU
        DEVICE_DATA_BUDGET_BYTES=120 << 30,
        # replay training steps from captured CUDA graphs once their input signature repeats
        USE_CUDA_GRAPHS=True,
        # data parallel only: split the guidance graph's edges over ranks by target-row blocks
        # (GraphConv forward = local block + all-gather); for graphs too large to be worth
        # aggregating redundantly on every rank
        PARTITION_GRAPH=False,
        # RMSprop as one fused kernel over flat parameter / gradient buffers
        USE_FUSED_OPTIMIZER=True,
        # run independent branches of a step (data encoders, graph branch) on forked streams
        MULTI_STREAM=True,
        # run the discriminator update beside the generator pass's encoders / decoders
        OVERLAP_DISCRIMINATOR=True,
        # fused kernels for the scalar tail of the objective (data-encoder KL, paired mean + cosine term)
        USE_FUSED_OBJECTIVE=True,
        # host-resident datasets: sparse count matrices travel to the device as (position, value) pairs
        # and are densified there, instead of being densified on the host and copied as [B, F] fp32
        SPARSE_HOST_BATCHES=True,
        # the decoder loss terms of one modality (reconstruction, joint cross, real cross) as one fused
        # launch sequence that shares the staged feature table and the observations
        USE_FUSED_TERMS=True,
        # activation + BatchNorm + dropout of the data encoders' hidden blocks in one launch each way
        # (the dropout masks then come from the kernel's own Philox stream, not from torch's)
        USE_FUSED_MLP=True,
        # the whole data-encoder MLP (Linear layers included) in h_depth + 1 launches each way
        # (training-mode minibatches of <= 128 cells, layer widths <= 256; needs USE_FUSED_MLP)
        USE_FUSED_ENCODER=True,
        # back-propagate the graph branch right behind the decoders (every gradient with respect to the
        # vertex embeddings is computed in the forward pass) instead of in the objective's backward pass
        EARLY_GRAPH_BACKWARD=True,
        # the generator's alignment term back-propagated through the discriminator right behind its forward pass
        EARLY_ALIGN_BACKWARD=True,
        # single-process loaders without pinned memory iterate their (torch) samplers directly instead of
        # through torch's DataLoader iterator: same index streams, a fraction of the per-minibatch host time
        FAST_LOADER=True,
        # negative sampling of a graph epoch as one native call (same MT19937 draws, no interpreter lock) instead
        # of the numpy restatement
        NATIVE_SAMPLER=True,
    )

    def __init__(self) -> None:
        for key, val in self._DEFAULTS.items():
            object.__setattr__(self, key, list(val) if isinstance(val, list) else val)
        # boolean switches from the environment, for A/B runs: GLUE_B200_FLAGS="EARLY_ALIGN_BACKWARD=0,FAST_LOADER=1"
        for item in filter(None, os.environ.get("GLUE_B200_FLAGS", "").split(",")):
            key, _, val = item.partition("=")
            if not isinstance(self._DEFAULTS.get(key.strip()), bool):
                raise AttributeError(f"GLUE_B200_FLAGS: {key!r} is not a boolean configuration key")
            object.__setattr__(self, key.strip(), bool(int(val)))

    def __setattr__(self, key: str, value: Any) -> None:
        if key not in self._DEFAULTS:
            raise AttributeError(f"Unknown configuration key {key!r}")
        if key == "CUDNN_MODE":
            if value not in ("repeatability", "performance"):
                raise ValueError("Invalid mode!")
            torch.backends.cudnn.deterministic = value == "repeatability"
            torch.backends.cudnn.benchmark = value == "performance"
        object.__setattr__(self, key, value)

    @property
    def DATALOADER_FETCHES_PER_BATCH(self) -> int:
        # scglue/utils.py:359-363
        return max(1, self.DATALOADER_NUM_WORKERS) * self.DATALOADER_FETCHES_PER_WORKER


config = _Settings()


class _LogManager:
    def __init__(self) -> None:
        self.console_log_level = logging.INFO
        self._handler = logging.StreamHandler()
        self._handler.setFormatter(logging.Formatter("[%(levelname)s] %(name)s: %(message)s"))

    def get_logger(self, name: str) -> logging.Logger:
        logger = logging.getLogger("glue_b200." + name)
        if not logger.handlers:
            logger.addHandler(self._handler)
            logger.setLevel(self.console_log_level)
            logger.propagate = False
        return logger


log = _LogManager()


def logged(obj):
    r"""Attach a ``logger`` attribute (class or function decorator)."""
    obj.logger = log.get_logger(obj.__name__)
    return obj


def get_rs(x: RandomState = None) -> np.random.RandomState:
    r"""Legacy ``RandomState`` from a seed / state / ``None`` (``scglue/utils.py:584-602``).
    The legacy MT19937 stream is what makes edge and cell sampling reproducible
    bit-for-bit across implementations."""
    if isinstance(x, int):
        return np.random.RandomState(x)
    if isinstance(x, np.random.RandomState):
        return x
    return np.random


class DelayedKeyboardInterrupt:  # pragma: no cover
    r"""Defer SIGINT until the block exits (``scglue/utils.py:458-481``)."""

    def __enter__(self):
        self._pending = None
        try:
            self._old = signal.signal(signal.SIGINT, self._catch)
        except ValueError:  # not in the main thread
            self._old = None
        return self

    def _catch(self, sig, frame):
        self._pending = (sig, frame)

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self._old is not None:
            signal.signal(signal.SIGINT, self._old)
            if self._pending:
                self._old(*self._pending)
