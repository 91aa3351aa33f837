"""Data-encoder MLP (fused kernels vs the per-op path): GPU time per forward + backward, replayed from a CUDA
graph; `ncu -k regex:mlp_ python tools/mlp_bench.py` profiles the fused kernels."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200.models import sc
from glue_b200.utils import config

dev = torch.device("cuda:0")
B, IN, H = 128, int(os.environ.get("IN", "100")), int(os.environ.get("H", "256"))
torch.manual_seed(0)
enc = sc.VanillaDataEncoder(IN, 50, h_depth=2, h_dim=H, dropout=0.2).to(dev)
xs = [torch.randn(B, IN, device=dev) for _ in range(4)]
wl, ws = torch.randn(B, 50, device=dev), torch.randn(B, 50, device=dev)
params = list(enc.parameters())


def step(i):
    u, _ = enc(xs[i % 4], xs[i % 4])
    loss = (u.mean * wl).sum() + (u.stddev * ws).sum()
    return torch.autograd.grad(loss, params)


def fwd_only(i):
    with torch.no_grad():
        u, _ = enc(xs[i % 4], xs[i % 4])
    return u.mean


def timed(fn, iters=8, replays=20):
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


out = {}
for fused in (True, False):
    config.USE_FUSED_ENCODER = fused
    tag = "fused" if fused else "per_op"
    out[f"{tag}_fwd_us"] = timed(fwd_only)
    out[f"{tag}_fwd_bwd_us"] = timed(step)
print(json.dumps(out, indent=1))
