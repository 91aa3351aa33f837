"""GPU-time micro-benchmark of the fused decoder (events recorded behind a blocker so that the
host launch path is off the critical path; distinct x batches cycle so inputs come from HBM).

    IMPLS=2,3 ONLY=atac ITERS=16 python tools/dec_bench.py

Per shape: one term (fwd+bwd and loss-only) per pinned implementation (``glue_decoder_args.impl``:
2 = per-row-staged tcgen05 kernels, 3 = TMA-staged), and the three paired terms (reconstruction,
joint cross, real cross) as one fused launch sequence vs. three single-term calls."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
B, K = 128, 50
ITERS = int(os.environ.get("ITERS", "16"))
IMPLS = [int(v) for v in os.environ.get("IMPLS", "2,3").split(",")]
FAMILY = os.environ.get("FAMILY", "NB")
NB = int(os.environ.get("NBATCH", "1"))
out = {}


def timed(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(ITERS):
        fn(i)
    host_s = time.perf_counter() - t0  # enqueue time of the timed calls
    torch.cuda.synchronize()
    # a spinning blocker the launches below queue behind, longer than their enqueue time (a GEMM blocker
    # would pull the SM clock down under the power cap and inflate what is timed behind it)
    torch.cuda._sleep(max(12_000_000, int(host_s * 1.5 * 2.0e9)))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(ITERS):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / ITERS


for name, F, rate in (("atac", 50000, 0.05), ("rna", 2000, 0.3), ("atac150k", 150000, 0.05)):
    if os.environ.get("ONLY") and os.environ["ONLY"] != name:
        continue
    us = [(torch.randn(B, K, device=dev) * 0.5).requires_grad_(True) for _ in range(3)]
    v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
    xs = [torch.poisson(torch.full((B, F), rate, device=dev)) for _ in range(8)]
    if os.environ.get("XDIST") == "bench":  # bench.py's counts: non-zeros at `rate`, values 1 + Poisson(0.3)
        xs = [(torch.rand(B, F, device=dev) < rate).float() * (1 + torch.poisson(torch.full((B, F), 0.3, device=dev)))
              for _ in range(8)]
    ls = [x.sum(1, keepdim=True) + 1 for x in xs]
    npar = 4 if FAMILY in ("ZINB", "ZILN") else 3
    ps = [torch.zeros(NB, F, device=dev, requires_grad=True) for _ in range(npar)]
    zi = ps[3] if npar == 4 else None
    bidx = torch.randint(0, NB, (B,), device=dev)
    rws = [None] + [torch.rand(B, device=dev) / (B * F) for _ in range(2)]
    coefs = [1.0, 0.02, 0.02]
    need_l = FAMILY in ("NB", "ZINB")
    P = 2 * npar

    def single(i, grad=True, t=0):
        kw = dict(b=bidx, l=ls[i % 8] if need_l else None, x=xs[i % 8], reduce="mean", row_weight=rws[t],
                  upstream=coefs[t])
        if grad:
            return ops.decoder_nll(FAMILY, us[t], v, ps[0], ps[1], ps[2], zi, **kw)
        with torch.no_grad():
            return ops.decoder_nll(FAMILY, us[t], v, ps[0], ps[1], ps[2], zi, **kw)

    def fused(i):
        return ops.decoder_nll_multi(FAMILY, us, v, ps[0], ps[1], ps[2], zi, b=bidx, l=ls[i % 8] if need_l else None,
                                     x=xs[i % 8], reduce="mean", row_weights=rws, upstreams=coefs)

    alg1 = B * F * 4 + 2 * F * K * 4 + P * NB * F * 4 + 2 * B * K * 4 + 16 * B
    alg3 = 3 * B * F * 4 + 2 * F * K * 4 + P * NB * F * 4 + 3 * (2 * B * K * 4 + 16 * B)
    for impl in IMPLS:
        ops.DECODER_IMPL = impl
        t = timed(lambda i: single(i, True))
        out[f"{name}_fwd_bwd_impl{impl}"] = {"us": t, "GBps": alg1 / t / 1e3, "frac_of_6554.9": alg1 / t / 1e3 / 6554.9}
        t = timed(lambda i: single(i, False))
        out[f"{name}_fwd_impl{impl}"] = {"us": t}
        t = timed(lambda i: [single(i, True, k) for k in range(3)])
        out[f"{name}_three_calls_impl{impl}"] = {"us": t, "GBps": alg3 / t / 1e3, "frac_of_6554.9": alg3 / t / 1e3 / 6554.9}
    ops.DECODER_IMPL = 0
    t = timed(fused)
    out[f"{name}_three_terms_fused"] = {"us": t, "GBps": alg3 / t / 1e3, "frac_of_6554.9": alg3 / t / 1e3 / 6554.9}
print(json.dumps(out, indent=1))
