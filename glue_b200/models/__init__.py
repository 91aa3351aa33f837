r"""
Public entry points of the integration models (API of ``scglue/models/__init__.py``):
:func:`configure_dataset`, :func:`fit_SCGLUE`, :func:`load_model`, and the model / module
classes.
"""

import os
from typing import Any, Mapping, Optional

import networkx as nx
import numpy as np
import pandas as pd

from ..utils import config, logged
from . import data, glue, nn, plugins, prob, sc, scglue  # noqa: F401
from .base import Model
from .scglue import (PairedSCGLUEModel, SCGLUEModel, register_prob_model)  # noqa: F401

Kws = Optional[Mapping[str, Any]]


@logged
def configure_dataset(adata, prob_model: str, use_highly_variable: bool = True,
                      use_layer: Optional[str] = None, use_rep: Optional[str] = None,
                      use_batch: Optional[str] = None, use_cell_type: Optional[str] = None,
                      use_dsc_weight: Optional[str] = None, use_obs_names: bool = False,
                      nan_sparse: bool = False) -> None:
    r"""
    Record in ``adata.uns[config.ANNDATA_KEY]`` how a dataset enters training
    (``scglue/models/__init__.py:26-139``).

    ``prob_model`` selects the decoder family (``"NB"``, ``"ZINB"``, ``"Normal"``,
    ``"ZILN"``, ``"ZIN"``, ``"Bernoulli"`` or anything added via ``register_prob_model``);
    ``use_highly_variable`` restricts features to ``adata.var["highly_variable"]``;
    ``use_layer`` / ``use_rep`` pick the count layer and the encoder input representation;
    ``use_batch`` / ``use_cell_type`` / ``use_dsc_weight`` name ``adata.obs`` columns;
    ``use_obs_names`` pairs cells across datasets by name; ``nan_sparse`` treats missing
    sparse entries as NaN.
    """
    if config.ANNDATA_KEY in adata.uns:
        configure_dataset.logger.warning("`configure_dataset` has already been called. "
                                         "Previous configuration will be overwritten!")

    def obs_levels(column: str) -> np.ndarray:
        return pd.Index(adata.obs[column]).dropna().drop_duplicates().sort_values().to_numpy()

    def require(key: Optional[str], container, label: str) -> Optional[str]:
        if key and key not in container:
            raise ValueError(f"Invalid `{label}`!")
        return key or None

    cfg = {"prob_model": prob_model, "use_highly_variable": bool(use_highly_variable)}
    if use_highly_variable:
        if "highly_variable" not in adata.var:
            raise ValueError("Please mark highly variable features first!")
        cfg["features"] = adata.var.query("highly_variable").index.to_numpy().tolist()
    else:
        cfg["features"] = adata.var_names.to_numpy().tolist()
    cfg["use_layer"] = require(use_layer, adata.layers, "use_layer")
    cfg["use_rep"] = require(use_rep, adata.obsm, "use_rep")
    cfg["rep_dim"] = adata.obsm[use_rep].shape[1] if use_rep else None
    cfg["use_batch"] = require(use_batch, adata.obs, "use_batch")
    cfg["batches"] = obs_levels(use_batch) if use_batch else None
    cfg["use_cell_type"] = require(use_cell_type, adata.obs, "use_cell_type")
    cfg["cell_types"] = obs_levels(use_cell_type) if use_cell_type else None
    cfg["use_dsc_weight"] = require(use_dsc_weight, adata.obs, "use_dsc_weight")
    cfg["use_obs_names"] = use_obs_names
    cfg["nan_sparse"] = nan_sparse
    adata.uns[config.ANNDATA_KEY] = cfg


def load_model(fname: os.PathLike) -> Model:
    r"""Load a model saved with ``Model.save`` (``scglue/models/__init__.py:142-161``)."""
    return Model.load(fname)


@logged
def fit_SCGLUE(adatas: Mapping[str, Any], graph: nx.Graph, model: type = SCGLUEModel,
               skip_balance: bool = False, init_kws: Kws = None, compile_kws: Kws = None,
               fit_kws: Kws = None, balance_kws: Kws = None) -> SCGLUEModel:
    r"""
    Two-stage training (``scglue/models/__init__.py:165-273``): pretrain without adversarial
    alignment, estimate per-cell balancing weights from the pretrained embeddings, then
    fine-tune a fresh model initialised from the pretrained one with alignment on.
    """
    init_kws, compile_kws = dict(init_kws or {}), dict(compile_kws or {})
    fit_kws, balance_kws = dict(fit_kws or {}), dict(balance_kws or {})
    vertices = sorted(graph.nodes)

    fit_SCGLUE.logger.info("Pretraining SCGLUE model...")
    stage_kws = {**fit_kws, "align_burnin": np.inf, "safe_burnin": False}
    if "directory" in stage_kws:
        stage_kws["directory"] = os.path.join(stage_kws["directory"], "pretrain")
    pretrain = model(adatas, vertices, **{**init_kws, "shared_batches": False})
    pretrain.compile(**compile_kws)
    pretrain.fit(adatas, graph, **stage_kws)
    if "directory" in stage_kws:
        pretrain.save(os.path.join(stage_kws["directory"], "pretrain.dill"))

    if not skip_balance:
        from ..data import estimate_balancing_weight
        fit_SCGLUE.logger.info("Estimating balancing weight...")
        rep = f"X_{config.TMP_PREFIX}"
        for k, adata in adatas.items():
            adata.obsm[rep] = pretrain.encode_data(k, adata)
        use_batch = None
        if init_kws.get("shared_batches"):
            names = {a.uns[config.ANNDATA_KEY]["use_batch"] for a in adatas.values()}
            use_batch = names.pop() if len(names) == 1 else None
        estimate_balancing_weight(*adatas.values(), use_rep=rep, use_batch=use_batch,
                                  key_added="balancing_weight", **balance_kws)
        for adata in adatas.values():
            adata.uns[config.ANNDATA_KEY]["use_dsc_weight"] = "balancing_weight"
            del adata.obsm[rep]

    fit_SCGLUE.logger.info("Fine-tuning SCGLUE model...")
    stage_kws = dict(fit_kws)
    if "directory" in stage_kws:
        stage_kws["directory"] = os.path.join(stage_kws["directory"], "fine-tune")
    finetune = model(adatas, vertices, **init_kws)
    finetune.adopt_pretrained_model(pretrain)
    finetune.compile(**compile_kws)
    finetune.random_seed += 1  # a different data order than the pretraining stage
    finetune.fit(adatas, graph, **stage_kws)
    if "directory" in stage_kws:
        finetune.save(os.path.join(stage_kws["directory"], "fine-tune.dill"))
    return finetune
