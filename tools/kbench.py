"""Micro-benchmark of the native kernels at BASELINE config-1 shapes (CUDA events, L2 flushed)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops
from glue_b200 import num as sampling
from tests.util import random_graph

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    times = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e3)
    return float(np.median(times)), float(np.min(times))


res = {}
rs = np.random.RandomState(0)
V, K, B = 52000, 50, 128
eidx, ewt, esgn = random_graph(rs, V, 100000)
enorm = sampling.normalize_edges(eidx, ewt)
csr = ops.GraphCSR(*(torch.as_tensor(a).to(dev) for a in (eidx, enorm, esgn)), V)
x = torch.randn(V, K, device=dev)
E = eidx.shape[1]
gc_bytes = 2 * V * K * 4 + 8 * E + 4 * (V + 1)
t = timeit(lambda: ops._spmm(csr.fwd, x))
res["graphconv_fwd_us"] = t
res["graphconv_fwd_frac_of_6554"] = gc_bytes / (t[0] * 1e-6) / 6554.9e9
t = timeit(lambda: ops._spmm(csr.bwd, x))
res["graphconv_bwd_us"] = t

Emb = 36000
s_eidx = torch.randint(0, V, (2, Emb), device=dev)
s_ewt = (torch.rand(Emb, device=dev) < 1 / 11).float()
s_sgn = torch.ones(Emb, device=dev)
v = (torch.randn(V, K, device=dev) * 0.3).requires_grad_(True)
res["graph_nll_fwd_us"] = timeit(lambda: ops.graph_nll(v, s_eidx, s_sgn, s_ewt))


def gfb():
    l = ops.graph_nll(v, s_eidx, s_sgn, s_ewt)
    l.backward()
    v.grad = None


res["graph_nll_fwd_bwd_us"] = timeit(gfb)

for name, F in (("atac", 50000), ("rna", 2000)):
    u = (torch.randn(B, K, device=dev) * 0.5).requires_grad_(True)
    vf = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
    xx = torch.poisson(torch.full((B, F), 0.05 if name == "atac" else 0.3, device=dev))
    l = xx.sum(1, keepdim=True) + 1
    ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]

    def fb():
        ops.decoder_nll("NB", u, vf, ps[0], ps[1], ps[2], None, l=l, x=xx)

    def fwd_only():
        with torch.no_grad():
            ops.decoder_nll("NB", u.detach(), vf.detach(), ps[0].detach(), ps[1].detach(), ps[2].detach(), None, l=l, x=xx)

    t = timeit(fb)
    alg = B * F * 4 + 2 * F * K * 4 + 2 * 3 * F * 4 + 2 * B * K * 4 + 16 * B
    res[f"decoder_nb_{name}_fwd_bwd_us"] = t
    res[f"decoder_nb_{name}_frac_of_6554"] = alg / (t[0] * 1e-6) / 6554.9e9
    res[f"decoder_nb_{name}_fwd_us"] = timeit(fwd_only)
print(json.dumps(res, indent=1))
