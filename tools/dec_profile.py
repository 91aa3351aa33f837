"""Per-kernel GPU time of the fused decoder call sequence in steady state (kineto), and the gaps
between the kernels of a call.   F=50000 TERMS=3 python tools/dec_profile.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import ProfilerActivity, profile

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
B, K, F = 128, 50, int(os.environ.get("F", "50000"))
T = int(os.environ.get("TERMS", "1"))
ops.DECODER_IMPL = int(os.environ.get("IMPL", "0"))
us = [(torch.randn(B, K, device=dev) * 0.5).requires_grad_(True) for _ in range(T)]
v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
xs = [torch.poisson(torch.full((B, F), 0.05, device=dev)) for _ in range(8)]
ls = [x.sum(1, keepdim=True) + 1 for x in xs]
ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]
rws = [torch.rand(B, device=dev) / (B * F) for _ in range(T)]


def call(i):
    if T == 1:
        return ops.decoder_nll("NB", us[0], v, ps[0], ps[1], ps[2], None, l=ls[i % 8], x=xs[i % 8], reduce="mean",
                               row_weight=rws[0], upstream=1.0)
    return ops.decoder_nll_multi("NB", us, v, ps[0], ps[1], ps[2], None, l=ls[i % 8], x=xs[i % 8], reduce="mean",
                                 row_weights=rws, upstreams=[1.0] * T)


for i in range(4):
    call(i)
torch.cuda.synchronize()
N = 12
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    torch.cuda._sleep(12_000_000)  # a spinning blocker: the launches below queue behind it (a GEMM blocker
    # would pull the SM clock down under the power cap and inflate what is timed behind it)
    for i in range(N):
        call(i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm" not in e.name.lower()
       and "cutlass" not in e.name.lower() and "nvjet" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in evs:
    agg.setdefault(e.name[:60], []).append(e.time_range.end - e.time_range.start)
tot = 0.0
for k, v_ in agg.items():
    print(f"{k:62s} n={len(v_):3d} mean={sum(v_) / len(v_):8.1f} us")
    tot += sum(v_) / N
span = (evs[-1].time_range.end - evs[0].time_range.start) / N
print(f"kernel time per call {tot:.1f} us; wall span per call {span:.1f} us; gaps {span - tot:.1f} us")
