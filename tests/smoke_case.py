"""One small invocation of every kernel on the hot path, checked against the CPU oracle.
Used by ``__graft_entry__.smoke()``; not collected by pytest."""
import numpy as np
import torch

from oracle import sampling
from oracle import scglue_cpu as oc
from tests.util import random_graph, rel_err


def run(device: torch.device) -> None:
    from glue_b200 import ops

    rs = np.random.RandomState(0)
    V, E, K, B, F = 600, 3000, 50, 32, 300
    eidx, ewt, esgn = random_graph(rs, V, E)
    enorm = sampling.normalize_edges(eidx, ewt)
    vrepr = torch.randn(V, K) * 0.3
    t = lambda a: torch.as_tensor(a)

    # GraphConv forward + backward
    ref_in = vrepr.clone().requires_grad_(True)
    ref = oc.graph_conv(ref_in, t(eidx), t(enorm), t(esgn))
    (ref_g,) = torch.autograd.grad(ref.square().sum(), ref_in)
    x = vrepr.to(device).requires_grad_(True)
    csr = ops.GraphCSR(t(eidx).to(device), t(enorm).to(device), t(esgn).to(device), V)
    out = ops.graph_conv(x, csr)
    (g,) = torch.autograd.grad(out.square().sum(), x)
    assert rel_err(out, ref) < 1e-5 and rel_err(g, ref_g) < 1e-5, "GraphConv parity"

    # graph decoder
    s_eidx = rs.randint(0, V, (2, 1100)).astype(np.int64)
    s_ewt = (rs.rand(1100) < 0.1).astype(np.float32)
    s_sgn = np.ones(1100, dtype=np.float32)
    ref_v = ref.detach().clone().requires_grad_(True)
    ref_nll = oc.graph_nll(ref_v, t(s_eidx), t(s_sgn), t(s_ewt))
    (ref_gv,) = torch.autograd.grad(ref_nll, ref_v)
    dv = out.detach().clone().requires_grad_(True)
    nll = ops.graph_nll(dv, t(s_eidx).to(device), t(s_sgn).to(device), t(s_ewt).to(device))
    (gv,) = torch.autograd.grad(nll, dv)
    assert rel_err(nll, ref_nll) < 1e-4 and rel_err(gv, ref_gv) < 1e-4, "graph decoder parity"

    # fused NB decoder
    gen = torch.Generator().manual_seed(1)
    u = torch.randn(B, K, generator=gen) * 0.5
    counts = torch.poisson(torch.rand(B, F, generator=gen) * 2.0, generator=gen)
    l = counts.sum(1, keepdim=True) + 1
    b = torch.zeros(B, dtype=torch.int64)
    names = ("scale_lin", "bias", "log_theta")
    params = {n: torch.randn(1, F, generator=gen) * 0.3 for n in names}
    vidx = torch.randperm(V, generator=gen)[:F]
    cu, cv = u.clone().requires_grad_(True), ref.detach().clone().requires_grad_(True)
    st = {f"u2x.m.{n}": p.clone().requires_grad_(True) for n, p in params.items()}
    ref_loss = -oc.nb_log_prob(st, "u2x.m.", cu, cv[vidx], b, l, counts).mean()
    ref_grads = torch.autograd.grad(ref_loss, [cu, cv] + list(st.values()))
    du, dvt = u.to(device).requires_grad_(True), out.detach().clone().requires_grad_(True)
    dp = [params[n].to(device).requires_grad_(True) for n in names]
    loss = ops.decoder_nll("NB", du, dvt, dp[0], dp[1], dp[2], None, b=b.to(device), l=l.to(device),
                           x=counts.to(device), vidx=vidx.to(device))
    grads = torch.autograd.grad(loss, [du, dvt] + dp)
    assert rel_err(loss, ref_loss) < 1e-4, "NB decoder loss parity"
    for got, want in zip(grads, ref_grads):
        assert rel_err(got, want) < 1e-4, "NB decoder gradient parity"
    torch.cuda.synchronize()
