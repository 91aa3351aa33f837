import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glue_b200.utils import config
from glue_b200.models import data as gdata
cfg = dict(bench.WORKLOADS[os.environ.get("CONFIG", "cfg2")])
adatas, graph = bench.make_workload(cfg)
dev = torch.device("cuda:0")
calls = []
orig = gdata.GraphDataset._propose_shuffle_native
def timed(self, seed):
    t = time.perf_counter(); out = orig(self, seed); calls.append(time.perf_counter() - t); return out
gdata.GraphDataset._propose_shuffle_native = timed
print("torch threads", torch.get_num_threads(), flush=True)
for workers, nthreads in ((4, 16),) * 10:
    config.GRAPH_SHUFFLE_NUM_WORKERS = workers
    torch.set_num_threads(nthreads)
    calls.clear(); a0 = gdata._PinnedSlabs.allocated
    r = bench.measure_fit(cfg, adatas, graph, dev)
    print(json.dumps({"graph_workers": workers, "torch_threads": nthreads, "cells_per_s": round(r["value"]), "s_per_epoch": round(r["s_per_epoch"], 4), "epoch_s": r["epoch_s"],
                      "slabs_pinned": gdata._PinnedSlabs.allocated - a0, "sampler_calls": len(calls),
                      "sampler_ms_mean": round(1e3 * sum(calls) / max(len(calls), 1), 1), "sampler_ms_max": round(1e3 * max(calls), 1)}), flush=True)
