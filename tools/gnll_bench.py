"""GPU-time micro-benchmark of graph_nll (fused-finalize forward + incidence sort + segmented backward)
on a sampled edge minibatch of the PBMC-shaped guidance graph (events behind a spinning blocker)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
V, K, G, E = 52000, 50, 2000, int(os.environ.get("E", "35860"))
# gene--peak pairs: every sampled edge has a gene endpoint (hubs), 1/11 positives, self loops now and then
genes, peaks = rs.randint(0, G, E), rs.randint(G, V, E)
flip = rs.rand(E) < 0.5
eidx = np.stack([np.where(flip, genes, peaks), np.where(flip, peaks, genes)]).astype(np.int64)
loops = rs.rand(E) < 0.3
eidx[1, loops] = eidx[0, loops]
s_eidx = torch.as_tensor(eidx, device=dev)
s_ewt = (torch.rand(E, device=dev) < 1 / 11).float()
s_sgn = torch.ones(E, device=dev)
vs = [(torch.randn(V, K, device=dev) * 0.3).requires_grad_(True) for _ in range(8)]
out = {}


def timed(fn, iters=8, replays=12):
    """GPU time per call: `iters` calls captured in a CUDA graph (as the training step replays them), the
    graph replayed back to back -- the Python launch path is out of the picture."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                fn(i)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (iters * replays)


def fwd(i):
    with torch.no_grad():
        ops.graph_nll(vs[i % 8], s_eidx, s_sgn, s_ewt)


def fwd_bwd(i):
    v = vs[i % 8]
    ops.graph_nll(v, s_eidx, s_sgn, s_ewt).backward()
    v.grad = None


touched = int(torch.unique(s_eidx).numel())
alg = E * 16 + min(2 * E, V) * K * 4 + V * K * 4
out["graph_nll_fwd_us"] = timed(fwd)
t = timed(fwd_bwd)
out["graph_nll_fwd_bwd_us"] = t
out["algorithmic_bytes"] = alg
out["frac_of_6554.9"] = alg / t / 1e3 / 6554.9
out["touched_vertices"] = touched
print(json.dumps(out, indent=1))

if os.environ.get("PROFILE"):
    import collections
    from torch.profiler import ProfilerActivity, profile

    for name, fn in (("fwd", fwd), ("fwd_bwd", fwd_bwd)):
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for i in range(8):
                fn(i)
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        agg = collections.OrderedDict()
        for e in sorted(evs, key=lambda e: e.time_range.start):
            agg.setdefault(e.name[:70], []).append(e.time_range.end - e.time_range.start)
        print(name, file=sys.stderr)
        for k, v_ in agg.items():
            print(f"  {k:72s} n={len(v_):3d} mean={sum(v_) / len(v_):8.1f} us", file=sys.stderr)
