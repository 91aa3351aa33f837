"""GPU-time micro-benchmark of the fused decoder (events recorded behind a blocker so that the
host launch path is off the critical path; distinct x batches cycle so inputs come from HBM)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
B, K = 128, 50
ITERS = int(os.environ.get("ITERS", "16"))
out = {}
blocker = torch.randn(8192, 8192, device=dev)
for name, F, rate in (("atac", 50000, 0.05), ("rna", 2000, 0.3), ("atac150k", 150000, 0.05)):
    if os.environ.get("ONLY") and os.environ["ONLY"] != name:
        continue
    u = (torch.randn(B, K, device=dev) * 0.5).requires_grad_(True)
    v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
    xs = [torch.poisson(torch.full((B, F), rate, device=dev)) for _ in range(8)]
    ls = [x.sum(1, keepdim=True) + 1 for x in xs]
    ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]

    def call(i, grad=True):
        if grad:
            return ops.decoder_nll("NB", u, v, ps[0], ps[1], ps[2], None, l=ls[i % 8], x=xs[i % 8])
        with torch.no_grad():
            return ops.decoder_nll("NB", u.detach(), v.detach(), ps[0].detach(), ps[1].detach(),
                                   ps[2].detach(), None, l=ls[i % 8], x=xs[i % 8])

    for variant, grad in [(vv, gg) for vv in os.environ.get("VARIANTS", "").split(",") for gg in (True, False)]:
        if "=" in variant:  # e.g. VARIANTS=GLUE_TC_PDL=0,GLUE_TC_PDL=1
            os.environ[variant.split("=")[0]] = variant.split("=")[1]
        elif variant:
            os.environ["GLUE_TC_VARIANT"] = variant  # read by the library at every call
        for i in range(3):
            call(i, grad)
        torch.cuda.synchronize()
        for _ in range(6):
            blocker @ blocker  # keeps the GPU busy while the launches below are queued
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(ITERS):
            call(i, grad)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / ITERS
        alg = B * F * 4 + (2 if grad else 1) * (F * K * 4 + 3 * F * 4 + B * K * 4) + 16 * B
        out[f"{name}_{'fwd_bwd' if grad else 'fwd'}" + (f"_variant{variant}" if variant else "")] = {"us": us, "GBps": alg / us / 1e3,
                                                         "frac_of_6554.9": alg / us / 1e3 / 6554.9}
print(json.dumps(out, indent=1))
