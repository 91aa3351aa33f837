r"""
Cell-level utilities that ``fit_SCGLUE`` needs between its two training stages
(``scglue/data.py:705-836``).

``estimate_balancing_weight`` has two halves in the reference: (1) cluster every dataset --
``scanpy.pp.neighbors`` (cosine kNN graph) + ``scanpy.tl.leiden`` -- and (2) match the clusters across
datasets and turn the match mass into per-cell weights.  Half (2) is restated here exactly (pinned by
``tests/test_data.py`` against ``oracle/balancing.py``, a line-by-line numpy restatement of
``scglue/data.py:783-836``).  Half (1) depends on scanpy / leidenalg / igraph, none of which is in this
image: **documented divergence** -- the built-in clustering is spherical k-means on the same cosine
geometry; callers who have scanpy can cluster with leiden themselves and pass the labels through
``use_clusters`` (an ``obs`` column), which makes the whole function the reference's.
"""

from typing import Optional

import numpy as np
import pandas as pd

from .utils import logged


def _cluster(rep: np.ndarray, resolution: float, seed: int = 0) -> np.ndarray:
    r"""Cluster cells on the cosine geometry of their embeddings.  The reference uses
    scanpy neighbours + leiden, neither of which exists in this image; spherical k-means
    with ``k ~ resolution * sqrt(n / 2)`` plays the same role (coarse, balanced groups)."""
    from sklearn.cluster import MiniBatchKMeans
    unit = rep / np.maximum(np.linalg.norm(rep, axis=1, keepdims=True), 1e-12)
    k = int(np.clip(round(resolution * np.sqrt(rep.shape[0] / 2)), 2, 50))
    k = min(k, rep.shape[0])
    return MiniBatchKMeans(n_clusters=k, random_state=seed, n_init=3).fit_predict(unit)


@logged
def estimate_balancing_weight(*adatas, use_rep: str = None, use_batch: Optional[str] = None,
                              resolution: float = 1.0, cutoff: float = 0.5, power: float = 4.0,
                              key_added: str = "balancing_weight",
                              use_clusters: Optional[str] = None) -> None:
    r"""
    Per-cell weights that balance matched cell populations across datasets: cluster each
    dataset, match cluster centroids by thresholded cosine similarity raised to ``power``,
    and weight every cell by its cluster's total match mass divided by the cluster size;
    weights are normalised to mean 1 and stored in ``adata.obs[key_added]``.

    ``use_clusters``: name of an ``obs`` column with precomputed cluster labels (e.g. leiden labels from
    scanpy, as the reference computes them); default: built-in spherical k-means (see module docstring).
    """
    if use_rep is None:
        raise ValueError("Missing required argument `use_rep`!")
    if use_batch:
        levels = [sorted(pd.unique(a.obs[use_batch])) for a in adatas]
        if any(lv != levels[0] for lv in levels):
            raise ValueError("Batches must match across datasets!")
        for adata in adatas:
            adata.obs[key_added] = np.nan
        for level in levels[0]:
            masks = [np.asarray(a.obs[use_batch] == level) for a in adatas]
            weights = _balance([np.asarray(a.obsm[use_rep])[m] for a, m in zip(adatas, masks)],
                               _labels(adatas, masks, use_rep, use_clusters, resolution), cutoff, power)
            for adata, mask, w in zip(adatas, masks, weights):
                col = adata.obs[key_added].to_numpy(copy=True)
                col[mask] = w
                adata.obs[key_added] = col
        return
    weights = _balance([np.asarray(a.obsm[use_rep]) for a in adatas],
                       _labels(adatas, [None] * len(adatas), use_rep, use_clusters, resolution), cutoff, power)
    for adata, w in zip(adatas, weights):
        adata.obs[key_added] = w


def _labels(adatas, masks, use_rep, use_clusters, resolution):
    r"""Integer cluster labels per dataset (restricted to ``masks``): the given ``obs`` column or the
    built-in clustering."""
    out = []
    for adata, mask in zip(adatas, masks):
        sel = slice(None) if mask is None else mask
        if use_clusters:
            if use_clusters not in adata.obs:
                raise ValueError(f"Cluster labels '{use_clusters}' cannot be found in input data!")
            out.append(pd.factorize(np.asarray(adata.obs[use_clusters])[sel], sort=True)[0])
        else:
            out.append(_cluster(np.asarray(adata.obsm[use_rep])[sel], resolution))
    return out


def _balance(reps, labels, cutoff, power):
    r"""``scglue/data.py:783-836`` given the cluster labels: l2-normalised cluster means, pairwise cosine
    with cutoff and power, product over dataset pairs, per-cluster match mass / cluster size, mean-1
    normalisation over cells."""
    cents, sizes = [], []
    for rep, lab in zip(reps, labels):
        k = lab.max() + 1
        size = np.bincount(lab, minlength=k).astype(float)
        cent = np.zeros((k, rep.shape[1]))
        np.add.at(cent, lab, rep)
        cent /= np.maximum(size[:, None], 1)
        cent /= np.maximum(np.linalg.norm(cent, axis=1, keepdims=True), 1e-12)
        cents.append(cent), sizes.append(size)
    n = len(reps)
    joint = np.ones([c.shape[0] for c in cents])
    for i in range(n):
        for j in range(i + 1, n):
            cos = cents[i] @ cents[j].T
            cos[cos < cutoff] = 0
            shape = [1] * n
            shape[i], shape[j] = cos.shape
            joint = joint * np.power(cos, power).reshape(shape)
    out = []
    for i in range(n):
        mass = joint.sum(axis=tuple(k for k in range(n) if k != i)) / np.maximum(sizes[i], 1)
        w = mass[labels[i]]
        total = w.sum()
        out.append(w / (total / w.size) if total > 0 else np.ones_like(w))
    return out
