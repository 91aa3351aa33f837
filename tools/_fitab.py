import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glue_b200.utils import config
cfg = dict(bench.WORKLOADS["cfg2"])
adatas, graph = bench.make_workload(cfg)
dev = torch.device("cuda:0")
for native, fast in ((True, True), (False, True), (True, False), (True, True)):
    config.NATIVE_SAMPLER, config.FAST_LOADER = native, fast
    r = bench.measure_fit(cfg, adatas, graph, dev)
    print(json.dumps({"native_sampler": native, "fast_loader": fast, "cells_per_s": r["value"], "s_per_epoch": r["s_per_epoch"],
                      "setup_s": r["setup_s"]}), flush=True)
