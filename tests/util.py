"""Helpers shared by the parity tests."""
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
