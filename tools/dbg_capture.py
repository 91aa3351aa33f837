import sys, os
sys.path.insert(0, '/root/repo')
os.environ["CUDA_LAUNCH_BLOCKING"] = "0"
import torch, numpy as np
from tests.test_step_gpu import build
g = dict(np.load('/root/repo/tests/golden/step_paired.npz'))
graph = dict(np.load('/root/repo/tests/golden/graph.npz'))
DEV='cuda:0'
keys=["rna","atac"]
B = g["x__rna"].shape[0]
t = lambda a: torch.as_tensor(a).to(DEV)
data = ([t(g[f"x__{k}"]) for k in keys] + [t(g[f"xrep__{k}"]) for k in keys] + [t(g[f"xbch__{k}"]) for k in keys]
        + [torch.full((B,), -1, dtype=torch.int64, device=DEV) for _ in keys] + [t(g[f"xdwt__{k}"]) for k in keys]
        + [t(g["pmsk"]), t(g["eidx"]), t(g["ewt"]), t(g["esgn"])])
class Eng:
    class state: epoch = int(g["epoch"])
net, trainer, _ = build(g, True)
trainer.eidx, trainer.enorm, trainer.esgn = (t(graph["eidx"]), t(graph["enorm_keepvar"]), t(graph["esgn"]))
trainer.align_burnin = int(g["align_burnin"])
for i in range(4):
    try:
        trainer.train_step(Eng, data)
    except Exception as e:
        print("step", i, "raised", repr(e)[:300]); break
    print("step", i, trainer.graphs.info())
print(getattr(trainer.graphs, "last_traceback", "no traceback"))
