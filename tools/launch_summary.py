"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share per kernel."""
import csv
import sys
from collections import defaultdict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "launch list")
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
rows = rows[skip:]
marker = sys.argv[4] if len(sys.argv) > 4 else None
if marker:  # one replayed step: the launches between the last two occurrences of the marker kernel
    at = [i for i, r in enumerate(rows) if marker in r["Kernel Name"]]
    if len(at) >= 2:
        rows = rows[at[-2]:at[-1]]
tot, agg = 0.0, defaultdict(lambda: [0, 0.0])
for r in rows:
    ns = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = ns / 1e3 if unit in ("ns", "nsecond") else (ns if unit.startswith("u") else ns * 1e3)
    name = r["Kernel Name"]
    agg[name][0] += 1
    agg[name][1] += us
    tot += us
print(f"# {title}\n")
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares).\n")
print("| share | total us | launches | us/launch | kernel |\n|---|---|---|---|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| {100 * us / tot:.1f}% | {us:.0f} | {n} | {us / n:.2f} | `{name[:110]}` |")
native = lambda name: "glue::" in name or "v3::" in name
ours = sum(us for name, (n, us) in agg.items() if native(name))
n_ours = sum(n for name, (n, us) in agg.items() if native(name))
dec = sum(us for name, (n, us) in agg.items() if "v3::" in name or "tc_" in name or "dec_" in name)
mlp = sum(us for name, (n, us) in agg.items() if "mlp_" in name)
print(f"\nKernels of libglue_b200.so (names under `glue::` / `v3::`): {n_ours} of the {len(rows)} launches, "
      f"**{100 * ours / tot:.1f} % of kernel time**; fused data-decoder kernels {100 * dec / tot:.1f} %, "
      f"encoder / discriminator MLP kernels {100 * mlp / tot:.1f} %.")
