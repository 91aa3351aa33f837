"""Helpers shared by the parity tests."""
import networkx as nx
import numpy as np
import torch


def rel_err(a, b) -> float:
    """max |a-b| / max |b|  (max-norm relative error; 0/0 -> 0)."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    denom = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / denom if denom > 0 else num


def assert_close(a, b, tol, what=""):
    err = rel_err(a, b)
    assert err <= tol, f"{what}: max-norm relative error {err:.3e} > {tol:.1e}"


def random_graph(rs, V, E, loops=True, neg_frac=0.2):
    """Random directed multi-graph triplets (int64 eidx, fp32 ewt, fp32 esgn)."""
    src = rs.randint(0, V, E)
    dst = rs.randint(0, V, E)
    if loops:
        src = np.concatenate([src, np.arange(V)])
        dst = np.concatenate([dst, np.arange(V)])
    eidx = np.stack([src, dst]).astype(np.int64)
    n = eidx.shape[1]
    ewt = rs.uniform(0.05, 1.0, n).astype(np.float32)
    esgn = np.where(rs.rand(n) < neg_frac, -1.0, 1.0).astype(np.float32)
    return eidx, ewt, esgn


def make_graph(rs, n_rna, n_atac, n_extra, n_pairs, neg_sign_frac=0.0):
    """Random bipartite guidance graph with self-loops, as an nx.Graph plus vertex list."""
    rna = [f"g{i}" for i in range(n_rna)]
    atac = [f"p{i}" for i in range(n_atac)]
    extra = [f"z{i}" for i in range(n_extra)]
    g = nx.Graph()
    g.add_nodes_from(rna + atac + extra)
    for _ in range(n_pairs):
        a, b = rna[rs.randint(n_rna)], atac[rs.randint(n_atac)]
        sign = -1 if rs.rand() < neg_sign_frac else 1
        g.add_edge(a, b, weight=float(rs.uniform(0.05, 1.0)), sign=sign)
    for e in extra[: n_extra // 2]:
        g.add_edge(e, rna[rs.randint(n_rna)], weight=float(rs.uniform(0.05, 1.0)), sign=1)
    for v in g.nodes:
        g.add_edge(v, v, weight=1.0, sign=1)
    vertices = sorted(g.nodes)
    return g, vertices, rna, atac


GOLDEN_GRAPH_ARGS = dict(seed=7, n_rna=12, n_atac=30, n_extra=6, n_pairs=70, neg_sign_frac=0.25)


def golden_graph():
    """The guidance graph tests/golden/graph.npz was generated from (same insertion order)."""
    a = dict(GOLDEN_GRAPH_ARGS)
    rs = np.random.RandomState(a.pop('seed'))
    return make_graph(rs, **a)


def assert_post_step_close(got, want, gref, tol, what=""):
    """Post-optimizer parameters.  RMSprop's first step is lr * g / (0.1 |g| + 1e-8): entries
    whose gradient is rounding noise (e.g. Linear biases in front of BatchNorm, whose true
    gradient is zero) take a full-size step in the direction of that noise -- in the
    reference as much as anywhere -- so they are compared only where the reference gradient
    is significant."""
    got, want = np.asarray(got), np.asarray(want)
    if gref is None:
        return assert_close(got, want, tol, what)
    live = np.abs(gref) > max(1e-6, 1e-3 * np.abs(gref).max())
    if live.any():
        assert_close(got[live], want[live], tol, what)
