"""CPU oracle for the integer / index side of the SCGLUE hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/scglue_cpu.py`` for the import rule).

numpy restatement of edge normalisation, CSR layout and per-graph-epoch negative
sampling.  Bit-exactness is the bar here, so the restatement keeps the
reference's accumulation order (sequential fp32, edge order) and its legacy
``numpy.random.RandomState`` call sequence.  Pinned by
``tests/test_oracle_golden.py`` against the reference's known-answer test
(``tests/test_num.py:76-113`` with the fixture of ``tests/conftest.py:277-289``)
and against ``GraphDataset.propose_shuffle`` outputs stored in ``tests/golden``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


# scglue/num.py:379-412 -- scipy COO row/column sums accumulate sequentially in
# the weight dtype; np.add.at reproduces that order exactly.
def vertex_degrees(eidx: np.ndarray, ewt: np.ndarray, vnum: Optional[int] = None,
                   direction: str = "both") -> np.ndarray:
    vnum = vnum or int(eidx.max()) + 1
    if direction not in ("in", "out", "both"):
        raise ValueError("Unrecognized direction!")

    def _acc(idx, w):
        deg = np.zeros(vnum, dtype=ewt.dtype)
        np.add.at(deg, idx, w)
        return deg

    if direction == "in":
        return _acc(eidx[1], ewt)
    if direction == "out":
        return _acc(eidx[0], ewt)
    loops = eidx[0] == eidx[1]
    return _acc(eidx[1], ewt) + _acc(eidx[0], ewt) - _acc(eidx[0][loops], ewt[loops])


# scglue/num.py:415-452
def normalize_edges(eidx: np.ndarray, ewt: np.ndarray, method: str = "keepvar") -> np.ndarray:
    if method not in ("in", "out", "sym", "keepvar"):
        raise ValueError("Unrecognized method!")
    enorm = ewt
    for side, row, methods in (("in", 1, ("in", "keepvar", "sym")), ("out", 0, ("out", "sym"))):
        if method in methods:
            deg = vertex_degrees(eidx, ewt, direction=side)
            with np.errstate(divide="ignore", invalid="ignore"):
                factor = np.power(deg[eidx[row]], -1 if method == side else -0.5)
            factor[~np.isfinite(factor)] = 0
            enorm = enorm * factor
    return enorm


# CSR definition of SURVEY.md section 7 ("bit-exact"): rows are edge targets,
# stable order inside a row, duplicates kept, val = (esgn * enorm) as nn.py:53.
def csr_by(key: np.ndarray, other: np.ndarray, val: np.ndarray, vnum: int
           ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    order = np.argsort(key, kind="stable")
    indptr = np.zeros(vnum + 1, dtype=np.int32)
    np.cumsum(np.bincount(key, minlength=vnum), out=indptr[1:])
    return indptr, other[order].astype(np.int32), val[order].astype(np.float32)


def csr_forward(eidx, enorm, esgn, vnum):
    return csr_by(eidx[1], eidx[0], (esgn * enorm).astype(np.float32), vnum)


def csr_transposed(eidx, enorm, esgn, vnum):
    return csr_by(eidx[0], eidx[1], (esgn * enorm).astype(np.float32), vnum)


class GraphSampler:
    """scglue/models/data.py:692-820 on ready-made triplets (the networkx ->
    triplet conversion of ``graph2triplet`` is host bookkeeping, not arithmetic)."""

    def __init__(self, eidx: np.ndarray, ewt: np.ndarray, esgn: np.ndarray, neg_samples: int = 1,
                 weighted_sampling: bool = True, deemphasize_loops: bool = True):
        self.eidx, self.ewt, self.esgn = eidx, ewt, esgn
        self.eset = {(i, j, s) for (i, j), s in zip(eidx.T, esgn)}
        self.vnum = int(eidx.max()) + 1
        if weighted_sampling:
            keep = eidx[0] != eidx[1] if deemphasize_loops else np.ones(ewt.size, dtype=bool)
            degree = vertex_degrees(eidx[:, keep], ewt[keep], vnum=self.vnum, direction="both")
        else:
            degree = np.ones(self.vnum, dtype=ewt.dtype)
        total = degree.sum()
        self.vprob = degree / total if total else np.ones(self.vnum, dtype=ewt.dtype) / self.vnum
        effective = ewt.sum()
        self.eprob = ewt / effective
        self.effective_enum = round(effective)
        self.neg_samples = neg_samples
        self.size = self.effective_enum * (1 + neg_samples)

    def propose_shuffle(self, seed: int):
        rs = np.random.RandomState(seed)
        pick = rs.choice(self.ewt.size, self.effective_enum, replace=True, p=self.eprob)
        pi, pj, ps = self.eidx[0][pick], self.eidx[1][pick], self.esgn[pick]
        pw = np.ones_like(self.ewt[pick])
        ni = np.tile(pi, self.neg_samples)
        ns = np.tile(ps, self.neg_samples)
        nw = np.zeros(pw.size * self.neg_samples, dtype=pw.dtype)
        nj = rs.choice(self.vnum, pj.size * self.neg_samples, replace=True, p=self.vprob)
        clash = np.where([t in self.eset for t in zip(ni, nj, ns)])[0]
        while clash.size:
            redraw = rs.choice(self.vnum, clash.size, replace=True, p=self.vprob)
            nj[clash] = redraw
            clash = clash[[t in self.eset for t in zip(ni[clash], redraw, ns[clash])]]
        idx = np.stack([np.concatenate([pi, ni]), np.concatenate([pj, nj])])
        w = np.concatenate([pw, nw])
        s = np.concatenate([ps, ns])
        perm = rs.permutation(idx.shape[1])
        return idx[:, perm], w[perm], s[perm]
