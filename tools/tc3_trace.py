"""Timeline of CTA 0 of the TMA-staged decoder kernels (impl bit 9: clock64 stamps in the workspace).

    F=50000 TERMS=3 python tools/tc3_trace.py

Prints, per kernel (0 = pass A, 1 = pass B, 2 = pass C) and unit, the stamps relative to the kernel's
entry, in cycles."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
B, K, F = 128, 50, int(os.environ.get("F", "50000"))
T = int(os.environ.get("TERMS", "1"))
us = [(torch.randn(B, K, device=dev) * 0.5).requires_grad_(True) for _ in range(T)]
v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
x = torch.poisson(torch.full((B, F), 0.05, device=dev))
if os.environ.get("XDIST") == "bench":  # bench.py's counts: 5 % non-zeros, values 1 + Poisson(0.3)
    x = (torch.rand(B, F, device=dev) < 0.05).float() * (1 + torch.poisson(torch.full((B, F), 0.3, device=dev)))
l = x.sum(1, keepdim=True) + 1
ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]
rws = [torch.rand(B, device=dev) / (B * F) for _ in range(T)]

captured = []
orig = ops._workspace


def grab(nbytes, device):
    ws = orig(nbytes, device)
    captured.append(ws)
    return ws


ops._workspace = grab
EV = {0: "tmaV", 1: "tmaU", 2: "mma_fwd", 3: "mma_xfull", 4: "mma_grad", 5: "epi_wait", 6: "epi_acc", 7: "epi_loop_end",
      8: "epi_arrived", 10: "cv_coef", 11: "cv_xempty", 12: "cv_done", 13: "ro_start", 14: "ro_done"}
KEV = {0: "entry", 1: "setup_done", 2: "epi_units_done", 3: "d2_ready", 4: "all_joined", 5: "end", 6: "ro0_begin",
       7: "ro0_rows", 8: "ro0_d1full", 9: "ro0_ld0", 10: "ro0_sts0", 11: "ro0_stg0", 12: "ro0_arrive"}
for rep in range(3):
    captured.clear()
    ops.DECODER_IMPL = 3 | 0x200
    losses = ops.decoder_nll_multi("NB", us, v, ps[0], ps[1], ps[2], None, l=l, x=x, reduce="mean", row_weights=rws,
                                   upstreams=[1.0] * T)
    torch.cuda.synchronize()
ws = captured[-1]
off = (-ws.data_ptr()) % 1024
raw = ws[off:off + 32768].view(torch.int64).cpu()
tr = raw[:768].reshape(3, 16, 16)
cta = raw[768:768 + 3 * 160 * 2].reshape(3, 160, 2)
epi = raw[1728:1728 + 3 * 160 * 3].reshape(3, 160, 3)
for k in (1, 2):
    live = cta[k][:, 0] > 0
    if live.any():
        dur = (cta[k][:, 1] - cta[k][:, 0])
        print(f"kernel {k}: per CTA (cta, smid, lifetime ns, epilogue loop cycles, logits-wait cycles), slowest 8 and fastest 4 of the 3-tile CTAs:")
        three = [c for c in range(160) if live[c] and c < (int(live.sum()) if False else 9999)]
        order = sorted(three, key=lambda c: -int(dur[c]))
        for c in order[:8] + order[60:64]:
            print(f"    cta {c:3d} smid {int(epi[k, c, 0]):3d}  {int(dur[c]):7d} ns  loop {int(epi[k, c, 1]):7d}  wait {int(epi[k, c, 2]):7d}")
for k in range(3):
    live = cta[k][cta[k, :, 0] > 0]
    if len(live):
        t_first = int(live[:, 0].min())
        dur = (live[:, 1] - live[:, 0])
        print(f"kernel {k}: {len(live)} CTAs; entry spread {int(live[:, 0].max()) - t_first} ns; lifetime min/mean/max "
              f"{int(dur.min())}/{int(dur.float().mean())}/{int(dur.max())} ns; last exit - first entry "
              f"{int(live[:, 1].max()) - t_first} ns; slowest CTAs {dur.argsort(descending=True)[:6].tolist()}")
for k in range(3):
    t0 = int(tr[k, 15, 0])
    if t0 == 0:
        continue
    if int(tr[k, 15, 13]):
        ns = int(tr[k, 15, 14]) - int(tr[k, 15, 13])
        cyc = int(tr[k, 15, 5]) - t0
        print(f"kernel {k}: CTA 0 lived {ns} ns = {cyc} cycles -> {cyc / max(ns, 1):.3f} GHz")
    print(f"kernel {k}: " + "  ".join(f"{KEV[e]}={int(tr[k, 15, e]) - t0}" for e in range(1, 13) if int(tr[k, 15, e])))
    for n in range(15):
        row = tr[k, n]
        if not int(row.abs().sum()):
            continue
        print(f"  unit {n:2d}: " + "  ".join(f"{EV[e]}={int(row[e]) - t0}" for e in sorted(EV) if int(row[e])))
