"""Kineto timeline of a few graph-replayed training steps (device-resident inputs): per-kernel
totals, per-stream busy time and the idle gaps of the whole device."""
import json, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

cfg = dict(bench.WORKLOADS[os.environ.get("CONFIG", "cfg2")])
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
# timeline of the last step: every activity of at least 5 us, and how many streams are busy over time
starts = [e for e in ev if "Memcpy DtoD" in e["name"] or "memcpy" in e["name"].lower()]
last0 = ev[len(ev) * 2 // 3]["ts"]
lastev = [e for e in ev if e["ts"] >= last0]
base = lastev[0]["ts"]
all_until = float(os.environ.get("ALL_UNTIL", "0"))  # list every activity that starts before this time (us)
all_from = float(os.environ.get("ALL_FROM", "0"))
if all_until:
    print(f"\nlast step, every activity starting in [{all_from:.0f}, {all_until:.0f}) us:")
    for e in lastev:
        if all_from <= e["ts"] - base < all_until:
            print(f"  {e['ts'] - base:7.1f} {e['dur']:6.1f}  s{e['args'].get('stream')}  {e['name'][:100]}")
print("\nlast step, activities >= 5 us (start us, dur us, stream, name):")
for e in lastev:
    if e["dur"] >= 5:
        print(f"{e['ts'] - base:8.0f} {e['dur']:7.1f}  s{e['args'].get('stream')}  {e['name'][:70]}")
span = max(e["ts"] + e["dur"] for e in lastev) - base
print("\nbusy streams per 50 us bucket:")
for b in range(0, int(span) + 1, 50):
    act = set()
    tot = 0.0
    for e in lastev:
        s0, s1 = e["ts"] - base, e["ts"] - base + e["dur"]
        ov = min(s1, b + 50) - max(s0, b)
        if ov > 0:
            act.add(e["args"].get("stream"))
            tot += ov
    print(f"{b:6d} us: {len(act)} streams, {tot / 50:.2f} average concurrency")
if os.environ.get("WINDOW"):
    w0, w1 = (float(v) for v in os.environ["WINDOW"].split(","))
    print(f"\nall activities of the last step starting in [{w0:.0f}, {w1:.0f}) us, in order:")
    prev_end = None
    for e in lastev:
        s0 = e["ts"] - base
        if w0 <= s0 < w1:
            gap = "" if prev_end is None else f"(+{s0 - prev_end:4.1f})"
            print(f"{s0:8.1f} {gap:8s} {e['dur']:6.1f}  s{e['args'].get('stream')}  {e['name'][:110]}")
            prev_end = s0 + e["dur"]
os.remove(path)
