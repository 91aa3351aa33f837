"""Small decoder / graph-encoder / optimizer invocations for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
from glue_b200 import ops
from glue_b200.optim import FlatRMSprop
from tests.test_ops_gpu import _decoder_case, _device_nll

for fam, B, F, K, nb in (("NB", 128, 300, 50, 1), ("NB", 77, 1300, 50, 1), ("ZINB", 100, 260, 50, 3),
                         ("NB", 300, 500, 50, 1), ("Normal", 130, 200, 16, 2), ("ZILN", 64, 129, 50, 1)):
    case = _decoder_case(fam, B, F, K, nb, seed=1)
    loss, grads = _device_nll(ops, fam, *case, "nanmean")
    torch.cuda.synchronize()
    print(fam, B, F, float(loss))
V, K = 1000, 50
ptr_ = torch.randn(V, K, device="cuda").requires_grad_(True)
wl = torch.randn(K, K, device="cuda").requires_grad_(True); ws = torch.randn(K, K, device="cuda").requires_grad_(True)
bl = torch.zeros(K, device="cuda").requires_grad_(True); bs = torch.zeros(K, device="cuda").requires_grad_(True)
vs, kl, mu, sg = ops.graph_encoder_head(ptr_, wl, bl, ws, bs, torch.randn(V, K, device="cuda"))
(vs.sum() + kl).backward()
ps = [torch.nn.Parameter(torch.randn(7, 3, device="cuda")), torch.nn.Parameter(torch.randn(1001, device="cuda"))]
opt = FlatRMSprop(ps, lr=1e-3)
for p in ps:
    p.grad = torch.randn_like(p)
opt.step()
torch.cuda.synchronize()
print("sanitize case done")
