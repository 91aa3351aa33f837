"""cProfile of a real ``model.fit`` (host side of the training loop): where the wall clock of an epoch goes.

    python tools/fit_profile.py            # cfg2, 1 warm-up fit + 4 profiled epochs
"""
import cProfile
import os
import pstats
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from glue_b200.utils import config

cfg = dict(bench.WORKLOADS[os.environ.get("CONFIG", "cfg2")])
adatas, graph = bench.make_workload(cfg)
model = bench.build_model(cfg, adatas, graph)
kw = dict(data_batch_size=cfg["batch"], patience=None, reduce_lr_patience=None)
config.PRINT_LOSS_INTERVAL = 1000
model.fit(adatas, graph, max_epochs=1, **kw)
torch.cuda.synchronize()
prof = cProfile.Profile()
prof.enable()
model.fit(adatas, graph, max_epochs=4, **kw)
torch.cuda.synchronize()
prof.disable()
st = pstats.Stats(prof)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
