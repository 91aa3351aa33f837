"""Pipeline timeline (clock64 stamps of CTA 0) of the tensor-core decoder kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from glue_b200 import ops, _lib

dev = torch.device("cuda:0")
B, K, F = 128, 50, int(os.environ.get("F", "50000"))
u = (torch.randn(B, K, device=dev) * 0.5).requires_grad_(True)
v = (torch.randn(F, K, device=dev) * 0.5).requires_grad_(True)
x = torch.poisson(torch.full((B, F), 0.05, device=dev))
l = x.sum(1, keepdim=True) + 1
ps = [torch.zeros(1, F, device=dev, requires_grad=True) for _ in range(3)]
for _ in range(3):
    ops.decoder_nll("NB", u, v, ps[0], ps[1], ps[2], None, l=l, x=x)
trace = torch.zeros(3 * 8 * 16, dtype=torch.int64, device=dev)
_lib.lib().glue_tc_set_trace(trace.data_ptr())
ops.decoder_nll("NB", u, v, ps[0], ps[1], ps[2], None, l=l, x=x)
torch.cuda.synchronize()
_lib.lib().glue_tc_set_trace(None)
tr = trace.cpu().view(3, 8, 16)
names = {0: "prod_start", 1: "prod_done", 2: "fwd_issued", 3: "mma_got_x_full", 4: "mma12_issued", 5: "epi_begin",
         6: "epi_acc_full", 7: "epi_x_empty", 8: "epi_loop_end_t0", 9: "epi_loop_end_t128", 10: "epi_after_bar",
         11: "epi_after_readout", 12: "prod_v_done", 13: "tail_d1_done",
         14: "tail_d2_full", 15: "tail_apart_done"}
for k, kn in enumerate(["rowstats", "main(B)", "softmax_bwd(C)"]):
    t0 = int(tr[k, 7, 2])
    print(f"== {kn}: entry->setup done {int(tr[k,7,0])-t0} cyc, total {int(tr[k,7,1])-t0} cyc")
    for it in range(7):
        ev = [(int(tr[k, it, e]) - t0, names[e]) for e in range(16) if int(tr[k, it, e]) != 0]
        if ev:
            print(f"  tile {it}: " + ", ".join(f"{n}={c}" for c, n in sorted(ev)))
