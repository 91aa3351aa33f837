r"""
Datasets, samplers and loaders (API of ``scglue/models/data.py``).

The index streams are the reference's, bit for bit: cell order comes from legacy
``numpy.random.RandomState(seed + epoch)`` permutations plus ``rs.choice`` random fill
for unpaired rows (``data.py:372-392,617-620``), edge minibatches from
``GraphDataset.propose_shuffle`` (``data.py:794-820``), and block-level shuffling is done
by the real ``torch.utils.data`` sampler classes with the same generator seeding
(``glue.py:580-625``).  What changes is where the bytes live: with ``to_device`` the count
matrices / representations / sampled edges stay resident in HBM and a minibatch is a
device-side row gather driven by a few hundred host-generated indices, instead of a host
densify + 26 MB H2D copy per step (``data.py:397-436``, ``scglue.py:276-289``).
"""

import copy
import functools
import operator
import threading
import uuid
from math import ceil
from typing import Any, Dict, List, Mapping, Optional, Tuple

import networkx as nx
import numpy as np
import pandas as pd
import scipy.sparse
import torch

from ..num import vertex_degrees
from ..utils import config, get_rs, logged
from .nn import get_default_numpy_dtype

DATA_CONFIG = Mapping[str, Any]


# --------------------------------------------------------------------------- base


@logged
class Dataset(torch.utils.data.Dataset):
    r"""
    Dataset with seed-driven custom shuffling.  ``shuffle()`` number ``n`` (0-based) of a
    dataset object uses seed ``random_seed + n`` (``data.py:99-112``); ``propose_shuffle`` is
    a pure function of the seed, so proposing the next shuffle ahead of time on a
    background thread (instead of the reference's worker process) cannot change results.
    """

    def __init__(self, getitem_size: int = 1) -> None:
        super().__init__()
        self.getitem_size = getitem_size
        self.shuffle_seed: Optional[int] = None
        self._lookahead: Dict[int, Tuple[threading.Thread, list]] = {}
        self._num_workers = 0

    def prepare_shuffle(self, num_workers: int = 1, random_seed: int = 0) -> None:
        r"""``num_workers`` background threads propose the shuffles of the next ``num_workers``
        seeds (numpy releases the GIL in the heavy parts).  A graph epoch of the PBMC shape is 32
        training steps = 35 ms of GPU time against 0.18 s to sample its edges, so one look-ahead
        is not enough to keep a B200 fed."""
        self.clean()
        self.shuffle_seed = random_seed
        self._num_workers = num_workers
        self._fill_lookahead()

    def _fill_lookahead(self) -> None:
        for seed in range(self.shuffle_seed, self.shuffle_seed + self._num_workers):
            if seed not in self._lookahead:
                self._spawn(seed)

    def _spawn(self, seed: int) -> None:
        box: list = []
        thread = threading.Thread(target=lambda: box.append(self.propose_shuffle(seed)),
                                  daemon=True)
        thread.start()
        self._lookahead[seed] = (thread, box)

    def shuffle(self) -> None:
        if self.shuffle_seed is None:
            raise RuntimeError("`prepare_shuffle` must be called before `shuffle`")
        ahead = self._lookahead.pop(self.shuffle_seed, None)
        if ahead is not None:
            thread, box = ahead
            thread.join()
            shuffled = box[0] if box else self.propose_shuffle(self.shuffle_seed)
        else:
            shuffled = self.propose_shuffle(self.shuffle_seed)
        self.accept_shuffle(shuffled)
        self.shuffle_seed += 1
        self._fill_lookahead()

    def propose_shuffle(self, seed: int) -> Any:
        raise NotImplementedError  # pragma: no cover

    def accept_shuffle(self, shuffled: Any) -> None:
        raise NotImplementedError  # pragma: no cover

    def clean(self) -> None:
        for thread, _ in (getattr(self, "_lookahead", None) or {}).values():
            thread.join()
        self._lookahead = {}

    @property
    def has_workers(self) -> bool:
        return bool(self._lookahead)


# ------------------------------------------------------------------------- arrays


def _rows(a, idx: np.ndarray) -> np.ndarray:
    block = a[idx]
    return block.toarray() if scipy.sparse.issparse(block) else np.asarray(block)


@logged
class ArrayDataset(Dataset):
    r"""Unpaired arrays; shorter arrays are recycled; sparse rows densified on fetch
    (``data.py:190-300``)."""

    def __init__(self, *arrays, getitem_size: int = 1) -> None:
        super().__init__(getitem_size=getitem_size)
        self.arrays = arrays

    @property
    def arrays(self):
        return self._arrays

    @arrays.setter
    def arrays(self, arrays) -> None:
        self.sizes = [a.shape[0] for a in arrays]
        if min(self.sizes) == 0:
            raise ValueError("Empty array is not allowed!")
        self.size = max(self.sizes)
        self.view_idx = [np.arange(s) for s in self.sizes]
        self.shuffle_idx = self.view_idx
        self._arrays = arrays

    def __len__(self) -> int:
        return ceil(self.size / self.getitem_size)

    def __getitem__(self, index: int) -> List[torch.Tensor]:
        pos = np.arange(index * self.getitem_size, min((index + 1) * self.getitem_size, self.size))
        return [torch.as_tensor(_rows(a, self.shuffle_idx[i][np.mod(pos, self.sizes[i])]))
                for i, a in enumerate(self.arrays)]

    def propose_shuffle(self, seed: int) -> List[np.ndarray]:
        rs = get_rs(seed)
        return [rs.permutation(v) for v in self.view_idx]

    def accept_shuffle(self, shuffled: List[np.ndarray]) -> None:
        self.shuffle_idx = shuffled

    def random_split(self, fractions: List[float], random_state=None) -> List["ArrayDataset"]:
        _check_fractions(fractions)
        rs = get_rs(random_state)
        cuts = np.cumsum(fractions)
        subs = [ArrayDataset(*self.arrays, getitem_size=self.getitem_size) for _ in fractions]
        for j, view in enumerate(self.view_idx):
            view = rs.permutation(view)
            pieces = np.split(view, np.round(cuts * view.size).astype(int)[:-1])
            for sub, piece in zip(subs, pieces):
                sub.sizes[j] = len(piece)
                sub.view_idx[j] = piece
                sub.shuffle_idx[j] = piece
        return subs


def _check_fractions(fractions: List[float]) -> None:
    if min(fractions) <= 0:
        raise ValueError("Fractions should be greater than 0!")
    if sum(fractions) != 1:
        raise ValueError("Fractions do not sum to 1!")


# ------------------------------------------------------------------------ AnnData


class DeviceCSR:
    r"""
    A sparse ``[N, F]`` matrix resident on the device as CSR, densified per minibatch with static
    shapes (no host read of the number of non-zeros): every requested row is read as
    ``max_row_nnz`` (column, value) slots, slots beyond the row's length repeat its last entry --
    duplicate writes of one value to one address -- and one ``scatter_`` fills the dense block.
    ``fill`` is the value of structural zeros (NaN for ``nan_sparse`` modalities, ``data.py:397-436``).
    Duck-types the one tensor method the loaders use: ``index_select(0, rows)``.
    """

    def __init__(self, csr, device: torch.device, fill: float = 0.0) -> None:
        csr = csr.tocsr()
        if not csr.has_canonical_format:  # (never touch the caller's matrix)
            csr = csr.copy()
            csr.sum_duplicates()
        self.shape = tuple(csr.shape)
        self.fill = fill
        self.indptr = torch.as_tensor(csr.indptr.astype(np.int64)).to(device)
        self.indices = torch.as_tensor(csr.indices.astype(np.int32)).to(device)
        self.data = torch.as_tensor(csr.data).to(device)
        self.max_row_nnz = int(np.diff(csr.indptr).max()) if csr.shape[0] else 0
        self._slots = torch.arange(max(self.max_row_nnz, 1), device=device).unsqueeze(0)

    @staticmethod
    def nbytes_of(csr) -> int:
        return int(csr.nnz) * (4 + csr.data.dtype.itemsize) + (csr.shape[0] + 1) * 8

    def index_select(self, dim: int, rows: torch.Tensor) -> torch.Tensor:
        if dim != 0:
            raise ValueError("DeviceCSR selects rows only")
        n, f = rows.shape[0], self.shape[1]
        out = torch.full((n, f), self.fill, dtype=self.data.dtype, device=self.data.device)
        if self.max_row_nnz == 0 or n == 0:
            return out
        start = self.indptr[rows]
        count = self.indptr[rows + 1] - start
        slot = torch.minimum(self._slots, (count - 1).clamp(min=0).unsqueeze(1))  # repeat the last entry
        pos = (start.unsqueeze(1) + slot).clamp(max=self.data.shape[0] - 1)
        empty = (count == 0).unsqueeze(1)
        cols = torch.where(empty, torch.zeros_like(pos), self.indices[pos].long())
        vals = torch.where(empty, out[:, :1], self.data[pos])  # rows without entries rewrite the fill
        return out.scatter_(1, cols, vals)


class SparseRows:
    r"""A block of rows of a sparse count matrix on its way to the device.  The reference densifies
    every minibatch on the host and copies ``[B, F]`` fp32 (``data.py:397-436``, 26 MB per step at the
    PBMC shape, > 90 % zeros); here the rows travel as (flat position, value) pairs -- about a tenth
    of the bytes -- and are scattered into a zeroed dense block on the device.  Eight ranks feeding
    from one host no longer saturate its memory / PCIe bandwidth."""

    BUCKET = 16384  # entries: padded lengths repeat, so device staging buffers are reused

    def __init__(self, csr) -> None:
        self.csr = csr.tocsr()

    @staticmethod
    def pack(blocks: List["SparseRows"]) -> torch.Tensor:
        r"""``int32 [2, cap]``: row 0 = flat positions ``row * F + col`` of the stacked blocks, row 1 =
        the fp32 values' bits; ``cap`` = the number of entries rounded up to ``BUCKET``, the padding
        repeats the last entry (the same value written to the same place again).  An all-zero batch
        carries one entry: 0 at position 0."""
        mat = scipy.sparse.vstack([b.csr for b in blocks], format="csr") if len(blocks) > 1 else blocks[0].csr
        n, f = mat.shape
        if n * f >= 2 ** 31:
            raise ValueError("minibatch too large for 32-bit flat positions")
        coo = mat.tocoo()
        flat = coo.row.astype(np.int64) * f + coo.col
        vals = coo.data.astype(np.float32, copy=False)
        if flat.size == 0:
            flat, vals = np.zeros(1, np.int64), np.zeros(1, np.float32)
        cap = -(-flat.size // SparseRows.BUCKET) * SparseRows.BUCKET
        out = np.empty((2, cap), dtype=np.int32)
        out[0, :flat.size], out[1, :flat.size] = flat, vals.view(np.int32)
        out[0, flat.size:], out[1, flat.size:] = flat[-1], vals.view(np.int32)[-1]
        return torch.as_tensor(out)


def is_packed_sparse(t: torch.Tensor) -> bool:
    return t.dtype == torch.int32 and t.dim() == 2 and t.shape[0] == 2


def unpack_sparse_rows(packed: torch.Tensor, n_rows: int, n_cols: int) -> torch.Tensor:
    r"""Dense fp32 ``[n_rows, n_cols]`` block of a :meth:`SparseRows.pack` ed minibatch, built where
    ``packed`` lives (three launches on the device: zero fill, index widening, scatter)."""
    out = torch.zeros(n_rows * n_cols, dtype=torch.float32, device=packed.device)
    out.scatter_(0, packed[0].long(), packed[1].view(torch.float32))
    return out.view(n_rows, n_cols)


@logged
class AnnDataset(Dataset):
    r"""
    Dataset over configured AnnData-like objects with partial pairing (``data.py:303-664``).

    A *view row* is one cell id of the union over all modalities; each view row maps to one
    row per modality (``shuffle_idx [n, n_modalities]``) plus a pairing mask
    (``shuffle_pmsk``) that is ``False`` where the modality does not contain that cell and a
    substitute row was filled in.

    Items are ``[x_k..., xrep_k..., xbch_k..., xlbl_k..., xdwt_k..., pmsk]`` like the
    reference.  After :meth:`to_device` the per-modality arrays are device tensors and an
    item is a device-side row gather.
    """

    def __init__(self, adatas, data_configs: List[DATA_CONFIG], mode: str = "train",
                 getitem_size: int = 1) -> None:
        super().__init__(getitem_size=getitem_size)
        if mode not in ("train", "eval"):
            raise ValueError("Invalid `mode`!")
        self.mode = mode
        self.device: Optional[torch.device] = None
        self.device_data = None
        # items carry sparse count rows as `SparseRows` (set by the training loaders for host-resident
        # datasets feeding a GPU); off: dense tensors, as the reference's dataset returns them
        self.sparse_items = False
        self.adatas = adatas
        self.data_configs = data_configs

    @property
    def adatas(self):
        return self._adatas

    @adatas.setter
    def adatas(self, adatas) -> None:
        self.sizes = [adata.shape[0] for adata in adatas]
        if min(self.sizes) == 0:
            raise ValueError("Empty dataset is not allowed!")
        self._adatas = adatas

    @property
    def data_configs(self):
        return self._data_configs

    @data_configs.setter
    def data_configs(self, data_configs) -> None:
        if len(data_configs) != len(self.adatas):
            raise ValueError("Number of data configs must match the number of datasets!")
        self.data_idx, self.extracted_data = self._extract_data(data_configs)
        self.view_idx = (pd.concat([idx.to_series() for idx in self.data_idx])
                         .drop_duplicates().to_numpy())
        self.size = self.view_idx.size
        self.shuffle_idx, self.shuffle_pmsk = self._get_idx_pmsk(self.view_idx)
        self.nan_sparse = [cfg["nan_sparse"] for cfg in data_configs]
        self._data_configs = data_configs

    # -- view rows -> per-modality rows ------------------------------------------------
    def _get_idx_pmsk(self, view_idx: Optional[np.ndarray], random_fill: bool = False,
                      random_state=None, indexers: Optional[List[np.ndarray]] = None
                      ) -> Tuple[np.ndarray, np.ndarray]:
        r"""Per-modality rows of the view rows ``view_idx`` (``indexers``: the ``get_indexer`` answers, when the
        caller has them already) and the mask of the cells each modality really has."""
        rs = get_rs(random_state) if random_fill else None
        cols, masks = [], []
        for m, data_idx in enumerate(self.data_idx):
            idx = data_idx.get_indexer(view_idx) if indexers is None else indexers[m]
            present = idx >= 0
            n_missing = present.size - present.sum()
            if random_fill:
                fill = rs.choice(idx[present], n_missing, replace=True)
            else:
                fill = idx[present][np.mod(np.arange(n_missing), present.sum())]
            idx[~present] = fill
            cols.append(idx)
            masks.append(present)
        return np.stack(cols, axis=1), np.stack(masks, axis=1)

    def __len__(self) -> int:
        return ceil(self.size / self.getitem_size)

    def __getitem__(self, index: int) -> List[torch.Tensor]:
        s = slice(index * self.getitem_size, min((index + 1) * self.getitem_size, self.size))
        if self.device is not None:  # indices only; gathered once per batch in `fetch_device`
            return [torch.as_tensor(self.shuffle_idx[s]), torch.as_tensor(self.shuffle_pmsk[s])]
        rows, pmsk = self.shuffle_idx[s].T, self.shuffle_pmsk[s]
        packed = self.sparse_items and self.mode == "train"
        items = []
        for g, group in enumerate(self.extracted_data):
            for idx, data, nan_sparse in zip(rows, group, self.nan_sparse):
                if packed and g == 0 and not nan_sparse and scipy.sparse.issparse(data):
                    items.append(SparseRows(data[idx]))  # densified on the device (`unpack_sparse_rows`)
                else:
                    items.append(torch.as_tensor(self._index_array(data, idx, nan_sparse)))
        items.append(torch.as_tensor(pmsk))
        return items

    @staticmethod
    def _index_array(arr, idx: np.ndarray, nan_sparse: bool) -> np.ndarray:
        arr = arr[idx]
        if scipy.sparse.issparse(arr):
            if nan_sparse:  # structural zeros mean "not measured"
                coo = arr.tocoo()
                dense = np.full(coo.shape, np.nan, dtype=coo.dtype)
                dense[coo.row, coo.col] = coo.data
                return dense
            return arr.toarray()
        return np.asarray(arr)

    # -- device residency ----------------------------------------------------------------
    def to_device(self, device: torch.device) -> "AnnDataset":
        r"""Upload the per-modality arrays once; minibatches become device row gathers.  Sparse
        matrices are densified if everything then fits ``config.DEVICE_DATA_BUDGET_BYTES`` (a
        minibatch is one ``index_select``); otherwise they stay CSR on the device
        (:class:`DeviceCSR`, ~20 x smaller at 5 % density) and rows are densified per minibatch."""
        device = torch.device(device)
        dense_bytes = sum(int(np.prod(arr.shape)) * 4 for group in self.extracted_data for arr in group)
        keep_sparse = dense_bytes > config.DEVICE_DATA_BUDGET_BYTES
        if keep_sparse:
            total = sum(DeviceCSR.nbytes_of(arr) if scipy.sparse.issparse(arr) else int(np.prod(arr.shape)) * 4
                        for group in self.extracted_data for arr in group)
            if total > config.DEVICE_DATA_BUDGET_BYTES:
                raise MemoryError(
                    f"device-resident data would need {total / 2**30:.1f} GiB even with the sparse "
                    f"matrices kept as CSR (> DEVICE_DATA_BUDGET_BYTES); keep the dataset on the host")
        dev_groups = []
        for group in self.extracted_data:
            dev_group = []
            for arr, nan_sparse in zip(group, self.nan_sparse):
                if scipy.sparse.issparse(arr) and keep_sparse:
                    dev_group.append(DeviceCSR(arr, device, fill=float("nan") if nan_sparse else 0.0))
                    continue
                dense = self._index_array(arr, slice(None), nan_sparse) \
                    if scipy.sparse.issparse(arr) else np.asarray(arr)
                dev_group.append(torch.as_tensor(dense).to(device))
            dev_groups.append(dev_group)
        self.device_data = dev_groups
        self.device = device
        return self

    def fetch_device(self, rows: torch.Tensor, pmsk: torch.Tensor) -> List[torch.Tensor]:
        r"""``rows [B, n_modalities]`` (host, int64) -> reference-format item on the device."""
        rows_dev = rows.to(self.device, non_blocking=True)
        items = [data.index_select(0, rows_dev[:, k])
                 for group in self.device_data for k, data in enumerate(group)]
        items.append(pmsk.to(self.device, non_blocking=True))
        return items

    # -- extraction (data.py:438-608) ----------------------------------------------------
    def _extract_data(self, data_configs):
        pairs = list(zip(self.adatas, data_configs))
        xuid = [self._extract_xuid(a, c) for a, c in pairs]
        xrep = [self._extract_xrep(a, c) for a, c in pairs]
        default = get_default_numpy_dtype()
        if self.mode == "eval":
            x = [np.empty((a.shape[0], 0), dtype=default) if r.size else self._extract_x(a, c)
                 for (a, c), r in zip(pairs, xrep)]
            xbch = xlbl = [np.empty((a.shape[0], 0), dtype=int) for a in self.adatas]
            xdwt = [np.empty((a.shape[0], 0), dtype=default) for a in self.adatas]
        else:
            x = [self._extract_x(a, c) for a, c in pairs]
            xbch = [self._extract_xbch(a, c) for a, c in pairs]
            xlbl = [self._extract_xlbl(a, c) for a, c in pairs]
            xdwt = [self._extract_xdwt(a, c) for a, c in pairs]
        return xuid, (x, xrep, xbch, xlbl, xdwt)

    @staticmethod
    def _extract_x(adata, cfg):
        features, use_layer = cfg["features"], cfg["use_layer"]
        if not np.array_equal(adata.var_names, features):
            adata = adata[:, features]
        if use_layer:
            if use_layer not in adata.layers:
                raise ValueError(f"Configured data layer '{use_layer}' "
                                 f"cannot be found in input data!")
            x = adata.layers[use_layer]
        else:
            x = adata.X
        default = get_default_numpy_dtype(np.iscomplexobj(x))
        if x.dtype.type is not default:
            if not (isinstance(x, np.ndarray) or scipy.sparse.issparse(x)):
                raise RuntimeError(f"User is responsible for ensuring a {default} dtype "
                                   f"when using backed data!")
            x = x.astype(default)
        return x.tocsr() if scipy.sparse.issparse(x) else x

    @staticmethod
    def _extract_xrep(adata, cfg):
        default = get_default_numpy_dtype()
        use_rep, rep_dim = cfg["use_rep"], cfg["rep_dim"]
        if not use_rep:
            return np.empty((adata.shape[0], 0), dtype=default)
        if use_rep not in adata.obsm:
            raise ValueError(f"Configured data representation '{use_rep}' "
                             f"cannot be found in input data!")
        xrep = np.asarray(adata.obsm[use_rep]).astype(default)
        if xrep.shape[1] != rep_dim:
            raise ValueError(f"Input representation dimensionality {xrep.shape[1]} "
                             f"does not match the configured {rep_dim}!")
        return xrep

    @staticmethod
    def _extract_xbch(adata, cfg):
        use_batch, batches = cfg["use_batch"], cfg["batches"]
        if not use_batch:
            return np.zeros(adata.shape[0], dtype=int)
        if use_batch not in adata.obs:
            raise ValueError(f"Configured data batch '{use_batch}' "
                             f"cannot be found in input data!")
        return pd.Index(batches).get_indexer(adata.obs[use_batch])

    @staticmethod
    def _extract_xlbl(adata, cfg):
        use_cell_type, cell_types = cfg["use_cell_type"], cfg["cell_types"]
        if not use_cell_type:
            return -np.ones(adata.shape[0], dtype=int)
        if use_cell_type not in adata.obs:
            raise ValueError(f"Configured cell type '{use_cell_type}' "
                             f"cannot be found in input data!")
        return pd.Index(cell_types).get_indexer(adata.obs[use_cell_type])

    @staticmethod
    def _extract_xdwt(adata, cfg):
        default = get_default_numpy_dtype()
        use_dsc_weight = cfg["use_dsc_weight"]
        if not use_dsc_weight:
            return np.ones(adata.shape[0], dtype=default)
        if use_dsc_weight not in adata.obs:
            raise ValueError(f"Configured discriminator sample weight '{use_dsc_weight}' "
                             f"cannot be found in input data!")
        xdwt = np.asarray(adata.obs[use_dsc_weight]).astype(default)
        return xdwt / (xdwt.sum() / xdwt.size)

    @staticmethod
    def _extract_xuid(adata, cfg) -> pd.Index:
        if cfg["use_obs_names"]:
            xuid = np.asarray(adata.obs_names)
        else:  # unpaired: every cell gets an id that collides with nothing
            xuid = np.array([uuid.uuid4().hex for _ in range(adata.shape[0])])
        if len(set(xuid)) != xuid.size:
            raise ValueError("Non-unique cell ID!")
        return pd.Index(xuid)

    # -- shuffling ------------------------------------------------------------------------
    def _view_indexers(self) -> List[np.ndarray]:
        r"""``get_indexer`` of the view's cell ids in every modality, looked up once per view (a hash lookup of
        9 000 string ids per modality costs 1.4 ms: more than two training steps)."""
        cache = self.__dict__.get("_indexer_cache")
        if cache is None or cache[0] is not self.view_idx:
            cache = self._indexer_cache = (self.view_idx, [d.get_indexer(self.view_idx) for d in self.data_idx])
        return cache[1]

    def propose_shuffle(self, seed: int):
        # ``rs.permutation(self.view_idx)`` (``data.py:617-620``) shuffles a copy of the ids with the swaps that
        # ``rs.permutation(n)`` applies to ``arange(n)``: permuting the cached row numbers gives the same rows
        rs = get_rs(seed)
        perm = rs.permutation(self.view_idx.size)
        return self._get_idx_pmsk(None, random_fill=True, random_state=rs,
                                  indexers=[rows[perm] for rows in self._view_indexers()])

    def accept_shuffle(self, shuffled) -> None:
        self.shuffle_idx, self.shuffle_pmsk = shuffled

    def random_split(self, fractions: List[float], random_state=None) -> List["AnnDataset"]:
        _check_fractions(fractions)
        rs = get_rs(random_state)
        view = rs.permutation(self.view_idx)
        pieces = np.split(view, np.round(np.cumsum(fractions) * view.size).astype(int)[:-1])
        subs = []
        for piece in pieces:
            sub = copy.copy(self)
            # a split owns its shuffle state: no look-ahead threads / seeds shared with the parent
            sub._lookahead, sub.shuffle_seed, sub._num_workers = {}, None, 0
            sub.view_idx, sub.size = piece, piece.size
            sub.shuffle_idx, sub.shuffle_pmsk = sub._get_idx_pmsk(piece)
            subs.append(sub)
        return subs


# -------------------------------------------------------------------------- graph


class _CdfSampler:
    r"""
    ``RandomState.choice(n, size, replace=True, p=p)`` without the per-call validation of ``p``
    and with most binary searches answered from a table.  The legacy implementation
    (numpy ``mtrand.pyx``, frozen by NEP 19) is ``cdf = p.astype(float64).cumsum(); cdf /= cdf[-1];
    u = random_sample(size); cdf.searchsorted(u, side="right")``.  Here ``lut[b]`` holds the
    answer for ``u = b / M`` (``M`` a power of two, so ``floor(u M)`` is exact): where
    ``lut[b] == lut[b + 1]`` no entry of the cdf lies in the bucket of ``u`` and the answer is
    ``lut[b]``; the other draws (a few percent) take the binary search.  Same uniforms, same
    answers, ~5 x faster for a million draws over 52 000 vertices.
    """

    def __init__(self, p: np.ndarray) -> None:
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        self.cdf = cdf
        self.bits = bits = min(22, max(10, int(np.ceil(np.log2(cdf.size))) + 4))
        self.m = 1 << bits
        self.lut = np.searchsorted(cdf, np.arange(self.m + 1) / self.m, side="right").astype(np.int64)

    def choice(self, rs: np.random.RandomState, size: int) -> np.ndarray:
        u = rs.random_sample(size)
        if size < 4096:
            return self.cdf.searchsorted(u, side="right")
        bucket = (u * self.m).astype(np.int64)
        out = self.lut[bucket]
        open_ = np.flatnonzero(self.lut[bucket + 1] != out)
        if open_.size:
            out[open_] = self.cdf.searchsorted(u[open_], side="right")
        return out


class _KeyFilter:
    r"""One-byte-per-slot hash filter over a set of int64 keys: ``maybe(keys)`` is ``False`` only
    for keys that are certainly not in the set (no false negatives), so that the exact lookup runs
    on the few percent that pass instead of on every candidate negative edge."""

    MULT = np.uint64(0x9E3779B97F4A7C15)

    def __init__(self, keys: np.ndarray) -> None:
        self.bits = bits = min(26, max(12, int(np.ceil(np.log2(max(keys.size, 1)))) + 6))
        self.shift = np.uint64(64 - bits)
        self.table = np.zeros(1 << bits, dtype=np.uint8)
        self.table[self._slot(keys)] = 1

    def _slot(self, keys: np.ndarray) -> np.ndarray:
        with np.errstate(over="ignore"):
            return ((keys.astype(np.uint64) * self.MULT) >> self.shift).astype(np.int64)

    def maybe(self, keys: np.ndarray) -> np.ndarray:
        return self.table[self._slot(keys)] != 0


class _PinnedSlabs:
    r"""Process-wide pool of pinned host slabs for the sampled graph epochs.  ``cudaHostAlloc`` / ``cudaFreeHost``
    of the 27 MB a graph epoch of the PBMC shape takes cost milliseconds each and stall kernel launches while
    they run, so slabs are recycled across graph epochs, datasets and fits instead of going back to torch's
    host allocator (whose cache ``torch.cuda.empty_cache()`` at the end of every fit would free again)."""

    _free: Dict[Tuple[str, int], List[torch.Tensor]] = {}
    _lock = threading.Lock()
    allocated = 0  # slabs ever pinned by this process

    @classmethod
    def take(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        with cls._lock:
            free = cls._free.get((str(device), nbytes))
            if free:
                return free.pop()
        cls.allocated += 1
        with torch.cuda.device(device):  # (a look-ahead thread starts on device 0: pin in the dataset's context)
            return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)

    @classmethod
    def give(cls, device: torch.device, slab: torch.Tensor) -> None:
        with cls._lock:
            cls._free.setdefault((str(device), slab.numel()), []).append(slab)


@logged
class GraphDataset(Dataset):
    r"""
    Guidance-graph dataset with negative sampling (``data.py:667-828``).

    Per graph epoch: draw ``round(sum(ewt))`` positive edges with replacement
    (``p = ewt / sum``), and for each ``neg_samples`` corrupted copies whose target is
    redrawn from the degree-weighted vertex distribution until ``(src, tgt, sign)`` is not a
    real edge; positives get weight 1, negatives 0; then permute.  The draw sequence on the
    legacy ``RandomState`` is the reference's, hence the edge stream is bit-identical; what
    changes is how the answers are computed: the membership test is vectorised (a one-byte hash
    filter in front of sorted int64 keys + ``searchsorted``, instead of a Python ``set`` of
    tuples), and ``RandomState.choice(n, size, p=p)`` -- uniforms followed by a binary search of
    the cumulative distribution -- keeps the uniforms but answers most of the searches from a
    bucket table (:class:`_CdfSampler`).  1.15 M edges per graph epoch at the PBMC shape: 0.63 s in
    the reference, 0.2 s here.
    """

    def __init__(self, graph: nx.Graph, vertices: pd.Index, neg_samples: int = 1,
                 weighted_sampling: bool = True, deemphasize_loops: bool = True,
                 getitem_size: int = 1) -> None:
        super().__init__(getitem_size=getitem_size)
        self.eidx, self.ewt, self.esgn = self.graph2triplet(graph, vertices)
        self.vnum = int(self.eidx.max()) + 1
        self._edge_keys = np.unique(self._key(self.eidx[0], self.eidx[1], self.esgn))
        if weighted_sampling:
            keep = (self.eidx[0] != self.eidx[1]) if deemphasize_loops \
                else np.ones(self.ewt.size, dtype=bool)
            degree = vertex_degrees(self.eidx[:, keep], self.ewt[keep], vnum=self.vnum,
                                    direction="both")
        else:
            degree = np.ones(self.vnum, dtype=self.ewt.dtype)
        total = degree.sum()
        self.vprob = degree / total if total \
            else np.ones(self.vnum, dtype=self.ewt.dtype) / self.vnum
        effective = self.ewt.sum()
        self.eprob = self.ewt / effective
        self._vsampler, self._esampler = _CdfSampler(self.vprob), _CdfSampler(self.eprob)
        self._edge_filter = _KeyFilter(self._edge_keys)
        self.effective_enum = round(effective)
        self.neg_samples = neg_samples
        self.size = self.effective_enum * (1 + self.neg_samples)
        self.samp_eidx = self.samp_ewt = self.samp_esgn = None
        self.device: Optional[torch.device] = None

    def _key(self, src, dst, sgn) -> np.ndarray:
        return (src.astype(np.int64) * self.vnum + dst.astype(np.int64)) * 2 + (sgn > 0)

    def _is_edge(self, src, dst, sgn) -> np.ndarray:
        keys = self._key(src, dst, sgn)
        hit = np.zeros(keys.shape, dtype=bool)
        maybe = np.flatnonzero(self._edge_filter.maybe(keys))  # no false negatives
        if maybe.size:
            cand = keys[maybe]
            pos = np.searchsorted(self._edge_keys, cand)
            pos[pos == self._edge_keys.size] = 0
            hit[maybe] = self._edge_keys[pos] == cand
        return hit

    @staticmethod
    def graph2triplet(graph: nx.Graph, vertices: pd.Index):
        r"""``(eidx int64 [2,E], ewt, esgn)`` in ``MultiDiGraph`` edge order; undirected input
        becomes bi-directed, multi-edges are kept (``data.py:733-779``)."""
        directed = nx.MultiDiGraph(graph)
        dtype = get_default_numpy_dtype()
        src, dst, weight, sign = [], [], [], []
        for (s, t, *_), attr in dict(directed.edges).items():
            src.append(s), dst.append(t)
            weight.append(attr["weight"]), sign.append(attr["sign"])
        vertices = pd.Index(vertices)
        eidx = np.stack([vertices.get_indexer(src), vertices.get_indexer(dst)]).astype(np.int64)
        if eidx.min() < 0:
            raise ValueError("Missing vertices!")
        ewt = np.asarray(weight).astype(dtype)
        if ewt.min() <= 0 or ewt.max() > 1:
            raise ValueError("Invalid edge weight!")
        esgn = np.asarray(sign).astype(dtype)
        if set(esgn).difference({-1, 1}):
            raise ValueError("Invalid edge sign!")
        return eidx, ewt, esgn

    def to_device(self, device: torch.device) -> "GraphDataset":
        self.device = torch.device(device)
        return self

    def __len__(self) -> int:
        return ceil(self.size / self.getitem_size)

    def __getitem__(self, index: int) -> List[torch.Tensor]:
        s = slice(index * self.getitem_size, min((index + 1) * self.getitem_size, self.size))
        return [self.samp_eidx[:, s], self.samp_ewt[s], self.samp_esgn[s]]

    def propose_shuffle(self, seed: int):
        r"""One graph epoch of sampled edges ``(eidx int64 [2, n], ewt [n], esgn [n])`` for ``seed``.  By
        default one native call (``glue_sample_graph_epoch``, ``csrc/sampler.cu``) that runs without the
        interpreter lock, so the look-ahead threads scale with the host cores; the numpy restatement below
        gives the same arrays bit for bit (``tests/test_data.py``) and takes over for a float64 default
        dtype."""
        if config.NATIVE_SAMPLER and self.esgn.dtype == np.float32 and self.size < 2 ** 32 \
                and isinstance(seed, (int, np.integer)) and 0 <= seed < 2 ** 32:
            return self._propose_shuffle_native(int(seed))
        return self._propose_shuffle_numpy(seed)

    def _native_sampler_tables(self):
        r"""Arrays ``glue_sample_graph_epoch`` reads (``include/glue_b200.h``), built once: contiguous edge
        triplets, 32-bit bucket tables of the two cumulative distributions (a quarter of a bucket per entry keeps
        them cache-resident) and the bit filter over the edge keys."""
        tables = self.__dict__.get("_native_tables")
        if tables is None:
            def lut(cdf):
                bits = min(22, max(10, int(np.ceil(np.log2(cdf.size))) + 2))
                grid = np.arange((1 << bits) + 1) / (1 << bits)
                return np.searchsorted(cdf, grid, side="right").astype(np.int32), bits
            keys = np.ascontiguousarray(self._edge_keys, dtype=np.int64)
            fbits = min(30, max(12, int(np.ceil(np.log2(max(keys.size, 1)))) + 5))
            with np.errstate(over="ignore"):
                slots = ((keys.astype(np.uint64) * _KeyFilter.MULT) >> np.uint64(64 - fbits)).astype(np.int64)
            flags = np.zeros(1 << fbits, dtype=np.uint8)
            flags[slots] = 1
            tables = self._native_tables = dict(
                esrc=np.ascontiguousarray(self.eidx[0], dtype=np.int64),
                edst=np.ascontiguousarray(self.eidx[1], dtype=np.int64),
                esgn=np.ascontiguousarray(self.esgn, dtype=np.float32), keys=keys,
                elut=lut(self._esampler.cdf), vlut=lut(self._vsampler.cdf),
                filter=(np.packbits(flags, bitorder="little"), fbits))
        return tables

    def _propose_shuffle_native(self, seed: int):
        from .. import _lib
        t = self._native_sampler_tables()
        n = self.size
        if self.device is not None and self.device.type == "cuda":
            # pinned, so that the upload of the epoch is asynchronous; the slab goes back to the pool once the
            # upload has run (`accept_shuffle`) or when the look-ahead is dropped (`clean`)
            slab = _PinnedSlabs.take(self.device, 24 * n)
            idx = slab[:16 * n].view(torch.int64).view(2, n)
            wt, sgn = slab[16 * n:20 * n].view(torch.float32), slab[20 * n:].view(torch.float32)
            self.__dict__.setdefault("_slabs", {})[idx.data_ptr()] = slab
        else:
            idx = torch.empty((2, n), dtype=torch.int64)
            wt, sgn = torch.empty(n, dtype=torch.float32), torch.empty(n, dtype=torch.float32)
        (elut, ebits), (vlut, vbits), (flt, fbits) = t["elut"], t["vlut"], t["filter"]
        _lib.check(_lib.lib().glue_sample_graph_epoch(
            seed, t["esrc"].ctypes.data, t["edst"].ctypes.data, t["esgn"].ctypes.data, t["esrc"].size,
            self._esampler.cdf.ctypes.data, elut.ctypes.data, ebits,
            self._vsampler.cdf.ctypes.data, vlut.ctypes.data, vbits, self.vnum,
            t["keys"].ctypes.data, t["keys"].size, flt.ctypes.data, fbits, self.effective_enum, self.neg_samples,
            idx[0].data_ptr(), idx[1].data_ptr(), wt.data_ptr(), sgn.data_ptr()), "sample_graph_epoch")
        return idx.numpy(), wt.numpy(), sgn.numpy()

    def _propose_shuffle_numpy(self, seed: int):
        rs = get_rs(seed)
        # (``_CdfSampler.choice`` == ``rs.choice(n, size, replace=True, p=p)``, draw for draw)
        pick = self._esampler.choice(rs, self.effective_enum)
        p_src, p_dst, p_sgn = self.eidx[0][pick], self.eidx[1][pick], self.esgn[pick]
        p_wt = np.ones_like(self.ewt[pick])
        n_src = np.tile(p_src, self.neg_samples)
        n_sgn = np.tile(p_sgn, self.neg_samples)
        n_wt = np.zeros(p_wt.size * self.neg_samples, dtype=p_wt.dtype)
        n_dst = self._vsampler.choice(rs, p_dst.size * self.neg_samples)
        clash = np.where(self._is_edge(n_src, n_dst, n_sgn))[0]
        while clash.size:  # NOTE: does not terminate on (near-)complete graphs, as in the reference
            redraw = self._vsampler.choice(rs, clash.size)
            n_dst[clash] = redraw
            clash = clash[self._is_edge(n_src[clash], redraw, n_sgn[clash])]
        wt, sgn = np.concatenate([p_wt, n_wt]), np.concatenate([p_sgn, n_sgn])
        perm = rs.permutation(wt.size)
        idx = np.empty((2, wt.size), dtype=np.int64)  # (row-wise takes: 4 x faster than idx[:, perm])
        np.take(np.concatenate([p_src, n_src]), perm, out=idx[0])
        np.take(np.concatenate([p_dst, n_dst]), perm, out=idx[1])
        return idx, wt[perm], sgn[perm]

    def accept_shuffle(self, shuffled) -> None:
        eidx, ewt, esgn = (torch.as_tensor(np.ascontiguousarray(a)) for a in shuffled)
        if self.device is not None:  # one upload per graph epoch; minibatches are slices
            eidx, ewt, esgn = (t.to(self.device, non_blocking=True) for t in (eidx, ewt, esgn))
            slab = self.__dict__.get("_slabs", {}).pop(shuffled[0].ctypes.data, None)
            if slab is not None:  # pinned by the native sampler: the upload is asynchronous
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                self._recycle_uploads(wait=False)
                self._uploads.append((done, slab))
        self.samp_eidx, self.samp_ewt, self.samp_esgn = eidx, ewt, esgn


    def _recycle_uploads(self, wait: bool) -> None:
        pending = []
        for done, slab in self.__dict__.get("_uploads", []):
            if wait:
                done.synchronize()
            if wait or done.query():
                _PinnedSlabs.give(self.device, slab)
            else:
                pending.append((done, slab))
        self._uploads = pending

    def clean(self) -> None:
        super().clean()  # (joins the look-ahead threads: nobody writes a slab any more)
        if self.__dict__.get("_uploads"):
            self._recycle_uploads(wait=True)
        for slab in self.__dict__.pop("_slabs", {}).values():  # proposed, never accepted
            _PinnedSlabs.give(self.device, slab)


# ------------------------------------------------------------------------ loaders


class DataLoader(torch.utils.data.DataLoader):
    r"""
    ``torch.utils.data.DataLoader`` that re-shuffles its dataset (new cell permutation / new
    negative samples) at every ``iter()`` and concatenates the fetched blocks
    (``data.py:831-859``).  For device-resident :class:`AnnDataset` s the blocks carry row
    indices only and the batch is gathered on the device here.
    """

    def __init__(self, dataset: Dataset, reshuffle: Optional[bool] = None, **kwargs) -> None:
        super().__init__(dataset, **kwargs)
        # `reshuffle` lets callers that pass an explicit `batch_sampler` keep the
        # re-shuffle-at-iter behaviour that `shuffle=True` implies
        self.shuffle = kwargs.get("shuffle", False) if reshuffle is None else reshuffle
        if isinstance(dataset, GraphDataset):
            self.collate_fn = self._collate_graph
        elif isinstance(dataset, AnnDataset) and dataset.device is not None:
            self.collate_fn = functools.partial(self._collate_device, dataset)
        else:
            self.collate_fn = self._collate

    def __iter__(self):
        if self.shuffle:
            self.dataset.shuffle()
        if (config.FAST_LOADER and self.num_workers == 0 and not self.pin_memory and self.timeout == 0
                and self.batch_sampler is not None and self.generator is not None):
            # Single-process iteration without torch's iterator machinery (0.3 ms per minibatch and loader:
            # more than the training step takes on the device).  Same index stream: the sampler objects are
            # torch's, and the one draw torch's iterator takes from the generator for its base seed -- after
            # creating the sampler iterator, before the first permutation -- is taken here too.
            sampler_iter = iter(self.batch_sampler)
            torch.empty((), dtype=torch.int64).random_(generator=self.generator)
            return self._fast_batches(sampler_iter)
        return super().__iter__()

    def _fast_batches(self, sampler_iter):
        fetch, collate = self.dataset.__getitem__, self.collate_fn
        for indices in sampler_iter:
            yield collate([fetch(i) for i in indices])

    @staticmethod
    def _collate(batch):
        return tuple(SparseRows.pack(list(parts)) if isinstance(parts[0], SparseRows) else torch.cat(parts, dim=0)
                     for parts in zip(*batch))

    @staticmethod
    def _collate_graph(batch):
        eidx, ewt, esgn = zip(*batch)
        return torch.cat(eidx, dim=1), torch.cat(ewt, dim=0), torch.cat(esgn, dim=0)

    @staticmethod
    def _collate_device(dataset: AnnDataset, batch):
        rows, pmsk = (torch.cat(parts, dim=0) for parts in zip(*batch))
        return tuple(dataset.fetch_device(rows, pmsk))


class ParallelDataLoader:
    r"""Iterate several loaders in lock-step, optionally cycling the shorter ones
    (``data.py:862-902``)."""

    def __init__(self, *data_loaders, cycle_flags: Optional[List[bool]] = None) -> None:
        cycle_flags = cycle_flags or [False] * len(data_loaders)
        if len(cycle_flags) != len(data_loaders):
            raise ValueError("Invalid cycle flags!")
        self.cycle_flags = cycle_flags
        self.data_loaders = list(data_loaders)
        self.num_loaders = len(self.data_loaders)
        self.iterators = None

    def __iter__(self) -> "ParallelDataLoader":
        self.iterators = [iter(loader) for loader in self.data_loaders]
        return self

    def _next(self, i: int):
        try:
            return next(self.iterators[i])
        except StopIteration:
            if not self.cycle_flags[i]:
                raise
            self.iterators[i] = iter(self.data_loaders[i])
            return next(self.iterators[i])

    def __next__(self):
        return functools.reduce(operator.add, [self._next(i) for i in range(self.num_loaders)])


class DevicePrefetcher:
    r"""
    Wrap a loader of host minibatches so that the host-to-device copy of minibatch ``i + 1``
    runs on a side stream while the training step of minibatch ``i`` executes (new design; the
    reference copies synchronously inside ``format_data``, ``scglue.py:276-289``).

    Host tensors should be pinned (``DataLoader(pin_memory=True)``) for the copies to be
    asynchronous.  Tensors that already live on ``device`` pass through.  Copies land in a ring
    of ``depth + 1`` pre-allocated device buffer sets.  Every yielded batch is safe to use on
    the current stream (it waits on the copy event) until the next ``depth`` batches have been
    requested: before a slot is overwritten the copy stream waits for everything the consumer
    stream had queued when the overwrite was requested.
    """

    def __init__(self, loader, device: torch.device, depth: int = 2) -> None:
        self.loader, self.device, self.depth = loader, torch.device(device), max(1, depth)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._rings, self._next_slot = {}, 0

    def __len__(self) -> int:
        return len(self.loader)

    def _buffers(self, batch, slot: int):
        r"""Device staging buffers of ring slot ``slot`` (allocated once per input signature: no
        allocator traffic, and no cross-stream frees, in the steady state)."""
        sig = tuple((tuple(t.shape), t.dtype) for t in batch)
        ring = self._rings.setdefault(sig, {})
        if slot not in ring:
            ring[slot] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch]
        return ring[slot]

    def _upload(self, batch):
        if self.stream is None:
            return list(batch), None
        slot = self._next_slot
        self._next_slot = (slot + 1) % (self.depth + 1)
        with torch.cuda.stream(self.stream):
            moved = []
            for buf, t in zip(self._buffers(batch, slot), batch):
                if t.device == self.device:
                    moved.append(t)
                else:
                    buf.copy_(t, non_blocking=True)
                    moved.append(buf)
            event = torch.cuda.Event()
            event.record(self.stream)
        return moved, event

    def __iter__(self):
        source = iter(self.loader)
        queue = []
        try:
            for _ in range(self.depth):
                queue.append(self._upload(next(source)))
        except StopIteration:
            pass
        while queue:
            moved, event = queue.pop(0)
            if event is not None:
                torch.cuda.current_stream(self.device).wait_event(event)
            try:
                if self.stream is not None:  # do not run ahead of the consumer by more than `depth`
                    self.stream.wait_stream(torch.cuda.current_stream(self.device))
                queue.append(self._upload(next(source)))
            except StopIteration:
                pass
            yield moved
