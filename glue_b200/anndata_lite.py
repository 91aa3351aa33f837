r"""
Minimal in-memory stand-in for :class:`anndata.AnnData`, used when the ``anndata`` package
is not installed (it is not in the B200 image).  It provides exactly what the model code
touches: ``X``, ``layers``, ``obs``, ``var``, ``obsm``, ``uns``, ``obs_names``,
``var_names``, ``shape`` and ``adata[:, features]`` / ``adata[rows]`` subsetting.  Real
``anndata.AnnData`` objects work everywhere this class does.
"""

from typing import Mapping, Optional

import numpy as np
import pandas as pd


class AnnData:
    def __init__(self, X=None, obs: Optional[pd.DataFrame] = None, var: Optional[pd.DataFrame] = None,
                 obsm: Optional[Mapping] = None, layers: Optional[Mapping] = None,
                 uns: Optional[Mapping] = None) -> None:
        n_obs = X.shape[0] if X is not None else len(obs)
        n_var = X.shape[1] if X is not None else (len(var) if var is not None else 0)
        self.X = X
        self.obs = obs if obs is not None else pd.DataFrame(index=[str(i) for i in range(n_obs)])
        self.var = var if var is not None else pd.DataFrame(index=[str(i) for i in range(n_var)])
        if len(self.obs) != n_obs or len(self.var) != n_var:
            raise ValueError("`obs` / `var` do not match the shape of `X`")
        self.obsm = dict(obsm or {})
        self.layers = dict(layers or {})
        self.uns = dict(uns or {})

    @property
    def shape(self):
        return (len(self.obs), len(self.var))

    @property
    def n_obs(self) -> int:
        return len(self.obs)

    @property
    def n_vars(self) -> int:
        return len(self.var)

    @property
    def obs_names(self) -> pd.Index:
        return self.obs.index

    @property
    def var_names(self) -> pd.Index:
        return self.var.index

    def _rows(self, key) -> np.ndarray:
        if isinstance(key, slice):
            return np.arange(self.n_obs)[key]
        key = np.asarray(key)
        if key.dtype == bool:
            return np.where(key)[0]
        if key.dtype.kind in "iu":
            return key
        return self.obs_names.get_indexer(key)

    def _cols(self, key) -> np.ndarray:
        if isinstance(key, slice):
            return np.arange(self.n_vars)[key]
        key = np.asarray(key)
        if key.dtype == bool:
            return np.where(key)[0]
        if key.dtype.kind in "iu":
            return key
        cols = self.var_names.get_indexer(key)
        if (cols < 0).any():
            raise KeyError("unknown variable names")
        return cols

    def __getitem__(self, key) -> "AnnData":
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        r, c = self._rows(rows), self._cols(cols)
        take = lambda m: None if m is None else m[r][:, c]
        return AnnData(X=take(self.X), obs=self.obs.iloc[r], var=self.var.iloc[c],
                       obsm={k: np.asarray(v)[r] for k, v in self.obsm.items()},
                       layers={k: take(v) for k, v in self.layers.items()}, uns=dict(self.uns))

    def copy(self) -> "AnnData":
        import copy
        return copy.deepcopy(self)
