"""Kineto timeline of a few graph-replayed training steps (device-resident inputs): per-kernel
totals, per-stream busy time and the idle gaps of the whole device."""
import json, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cfg = dict(bench.WORKLOAD)
adatas, graph = bench.make_workload(cfg)
model = bench.build_model(cfg, adatas, graph)
trainer = model.trainer
gds, loader, _ = bench.make_loaders(model, adatas, graph, cfg, device_resident=True)
trainer.set_full_graph(gds)
trainer.align_burnin = 57
it = iter(loader)
pool = [next(it) for _ in range(4)]
for i in range(6):
    trainer.train_step(bench._Engine, pool[i % 4])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        trainer.train_step(bench._Engine, pool[i % 4])
    torch.cuda.synchronize()
path = "gpurun_out/step_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} device activities over {(t1 - t0) / 3:.0f} us per step")
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in sorted(streams.items(), key=lambda kv: -sum(x["dur"] for x in kv[1])):
    print(f"stream {s}: {len(es) / 3:.0f} launches/step, busy {sum(x['dur'] for x in es) / 3:.0f} us/step")
# union busy time (any stream)
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, d = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, d
    else:
        cur_e = max(cur_e, d)
busy += cur_e - cur_s
print(f"device busy (union over streams): {busy / 3:.0f} us/step; idle {(t1 - t0 - busy) / 3:.0f} us/step")
agg = {}
for e in ev:
    k = e["name"][:90]
    n, d = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, d + e["dur"])
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{d / 3:8.1f} us/step  {n / 3:6.1f} x {d / n:7.2f} us  {k}")
os.remove(path)
