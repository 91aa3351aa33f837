r"""
Edge-level numerics shared by the graph encoder and the samplers
(host side, numpy; mirrors the part of ``scglue/num.py`` the hot path uses).
"""

from typing import Optional

import numpy as np

EPS = 1e-7  # scglue/num.py:12


def _scatter_sum(index: np.ndarray, weight: np.ndarray, size: int) -> np.ndarray:
    # Sequential accumulation in the weight dtype, edge order: this is the order in which the
    # reference's scipy COO row/column sums add (scglue/num.py:403-411), hence bit-identical.
    out = np.zeros(size, dtype=weight.dtype)
    np.add.at(out, index, weight)
    return out


def vertex_degrees(eidx: np.ndarray, ewt: np.ndarray, vnum: Optional[int] = None,
                   direction: str = "both") -> np.ndarray:
    r"""
    Weighted vertex degrees of an edge list (reference: ``scglue/num.py:379-412``).

    ``direction`` is one of ``"in"`` (sum over incoming edges, i.e. by ``eidx[1]``),
    ``"out"`` (by ``eidx[0]``) or ``"both"`` (in + out, self-loops counted once).
    """
    if direction not in ("in", "out", "both"):
        raise ValueError("Unrecognized direction!")
    eidx, ewt = np.asarray(eidx), np.asarray(ewt)
    vnum = vnum or int(eidx.max()) + 1
    if direction == "in":
        return _scatter_sum(eidx[1], ewt, vnum)
    if direction == "out":
        return _scatter_sum(eidx[0], ewt, vnum)
    loop = eidx[0] == eidx[1]
    return (_scatter_sum(eidx[1], ewt, vnum) + _scatter_sum(eidx[0], ewt, vnum)
            - _scatter_sum(eidx[0][loop], ewt[loop], vnum))


def normalize_edges(eidx: np.ndarray, ewt: np.ndarray, method: str = "keepvar") -> np.ndarray:
    r"""
    Normalised edge weights (reference: ``scglue/num.py:415-452``).

    ``"in"``: ``w / indeg[t]``; ``"out"``: ``w / outdeg[s]``; ``"sym"``:
    ``w / sqrt(indeg[t] outdeg[s])``; ``"keepvar"`` (what training uses,
    ``scglue/models/glue.py:555``): ``w / sqrt(indeg[t])``.  Edges into
    zero-degree vertices get weight 0.
    """
    if method not in ("in", "out", "sym", "keepvar"):
        raise ValueError("Unrecognized method!")
    eidx, ewt = np.asarray(eidx), np.asarray(ewt)
    power = {"in": (-1.0, None), "out": (None, -1.0), "sym": (-0.5, -0.5), "keepvar": (-0.5, None)}
    enorm = ewt
    for side, endpoint, exponent in zip(("in", "out"), (1, 0), power[method]):
        if exponent is None:
            continue
        degree = vertex_degrees(eidx, ewt, direction=side)[eidx[endpoint]]
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.power(degree, int(exponent) if exponent == -1.0 else exponent)
        factor[~np.isfinite(factor)] = 0
        enorm = enorm * factor
    return enorm
