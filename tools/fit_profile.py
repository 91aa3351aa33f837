"""cProfile of the steady state of a real ``model.fit`` (host side of the training loop): where the wall clock of
an epoch goes once datasets, graph capture and allocator are warm.

    python tools/fit_profile.py            # cfg2: a 1-epoch fit, then a 6-epoch fit whose epochs 2-6 are profiled
"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from glue_b200.models.base import EPOCH_COMPLETED, TrainingPlugin
from glue_b200.utils import config


class SteadyStateProfile(TrainingPlugin):
    """Turns the profiler on at the end of the first epoch; off when the fit returns."""

    def __init__(self):
        self.prof, self.tic, self.epochs = cProfile.Profile(), None, 0

    def attach(self, net, trainer, train_engine, val_engine, train_loader, val_loader, directory):
        def end(engine):
            torch.cuda.synchronize()
            if self.tic is None:
                self.tic = time.perf_counter()
                self.prof.enable()
            else:
                self.epochs += 1
                self.toc = time.perf_counter()
        train_engine.add_event_handler(EPOCH_COMPLETED, end)


cfg = dict(bench.WORKLOADS[os.environ.get("CONFIG", "cfg2")])
adatas, graph = bench.make_workload(cfg)
model = bench.build_model(cfg, adatas, graph)
kw = dict(data_batch_size=cfg["batch"], patience=None, reduce_lr_patience=None)
config.PRINT_LOSS_INTERVAL = 1000
model.fit(adatas, graph, max_epochs=1, **kw)
torch.cuda.synchronize()


def timed(owner, name, table):
    """Accumulate the wall clock of ``owner.name`` calls in ``table[label]`` (perf_counter, no profiler)."""
    fn = getattr(owner, name)
    label = f"{owner.__name__}.{name}"

    def wrapper(*a, **k):
        tic = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            cell = table.setdefault(label, [0, 0.0])
            cell[0] += 1
            cell[1] += time.perf_counter() - tic
    static = isinstance(owner.__dict__.get(name), staticmethod)
    setattr(owner, name, staticmethod(wrapper) if static else wrapper)
    return lambda: setattr(owner, name, staticmethod(fn) if static else fn)


# pass 1: timers around the pieces of the loop (host wall clock; the device runs asynchronously behind)
from glue_b200.models import data as gdata
table, undo = {}, []
for owner, name in ((gdata.ParallelDataLoader, "__next__"), (gdata.ParallelDataLoader, "__iter__"),
                    (gdata.GraphDataset, "shuffle"), (gdata.AnnDataset, "shuffle"),
                    (gdata.GraphDataset, "accept_shuffle"), (gdata.DataLoader, "_collate_graph"),
                    (gdata.AnnDataset, "fetch_device"), (type(model.trainer), "train_step"),
                    (type(model.trainer), "val_step")):
    undo.append(timed(owner, name, table))
clock = SteadyStateProfile()
clock.prof = type("NoProfile", (), {"enable": lambda self: None, "disable": lambda self: None})()
bench.run_fit(model, adatas, graph, cfg, kw, 6, clock)
for u in undo:
    u()
print(f"steady state, timers only: {(clock.toc - clock.tic) / clock.epochs:.4f} s per epoch; whole fit of 6 epochs:")
for label, (calls, secs) in sorted(table.items(), key=lambda kv: -kv[1][1]):
    print(f"  {label:42s} {calls:6d} calls  {secs * 1e3:9.1f} ms  {secs / calls * 1e6:9.1f} us / call")

# pass 2: cProfile of the same
plugin = SteadyStateProfile()
bench.run_fit(model, adatas, graph, cfg, kw, 6, plugin)
plugin.prof.disable()
print(f"steady state: {plugin.epochs} epochs, {(plugin.toc - plugin.tic) / plugin.epochs:.4f} s per epoch (profiled)")
st = pstats.Stats(plugin.prof)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
