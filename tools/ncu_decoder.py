"""The fused decoder call of the ATAC modality as the paired step issues it (three loss terms, B = 128,
F = 50 000) -- a short program for ncu:

    python tools/ncu_decoder.py                     # plain run first (must exit 0)
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/dec_traffic.csv python tools/ncu_decoder.py
    ncu --set full --clock-control none --import-source on -k regex:main3_kernel -s 6 -c 2 \
        -o gpurun_out/dec_main3 python tools/ncu_decoder.py

TERMS=1 gives the single-term call."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from glue_b200 import ops

dev = torch.device("cuda:0")
B, K, F = 128, 50, int(os.environ.get("F", "50000"))
T = int(os.environ.get("TERMS", "3"))
torch.manual_seed(0)
us = [(torch.randn(B, K, device=dev) * 0.5).requires_grad_(True) for _ in range(T)]
v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
xs = [torch.poisson(torch.full((B, F), 0.05, device=dev)) for _ in range(6)]
ls = [x.sum(1, keepdim=True) + 1 for x in xs]
ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]
rws = [None] + [torch.rand(B, device=dev) / (B * F) for _ in range(T - 1)]
for i in range(int(os.environ.get("CALLS", "4"))):
    if T == 1:
        loss = ops.decoder_nll("NB", us[0], v, ps[0], ps[1], ps[2], None, l=ls[i % 6], x=xs[i % 6], reduce="mean",
                               upstream=1.0)
    else:
        loss = ops.decoder_nll_multi("NB", us, v, ps[0], ps[1], ps[2], None, l=ls[i % 6], x=xs[i % 6], reduce="mean",
                                     row_weights=rws, upstreams=[1.0, 0.02, 0.02][:T])
torch.cuda.synchronize()
print("ok", [float(t) for t in (loss if T > 1 else [loss])])
