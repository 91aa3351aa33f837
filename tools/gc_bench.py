"""GPU-time micro-benchmark of the CSR GraphConv kernel (events behind a blocker)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
from glue_b200 import ops, num
from tests.util import random_graph
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
out = {}
for name, V, E in (("cfg1", 52000, 100000), ("atlas", 200000, 1000000)):
    eidx, ewt, esgn = random_graph(rs, V, E)
    enorm = num.normalize_edges(eidx, ewt)
    csr = ops.GraphCSR(*(torch.as_tensor(a).to(dev) for a in (eidx, enorm, esgn)), V)
    xs = [torch.randn(V, 50, device=dev) for _ in range(8)]
    for part, tag in ((csr.fwd, "fwd"), (csr.bwd, "bwd")):
        for i in range(3):
            ops._spmm(part, xs[i])
        torch.cuda.synchronize()
        torch.cuda._sleep(12_000_000)  # a spinning blocker: the launches below queue behind it (a GEMM blocker
        # would pull the SM clock down under the power cap and inflate what is timed behind it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(32):
            ops._spmm(part, xs[i % 8])
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 32
        nE = eidx.shape[1]
        alg = 2 * V * 50 * 4 + 8 * nE + 4 * (V + 1)
        out[f"{name}_{tag}"] = {"us": us, "GBps": alg / us / 1e3, "frac_of_6554.9": alg / us / 1e3 / 6554.9}
print(json.dumps(out, indent=1))
