"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python scripts/launch_breakdown.py gpurun_out/launches.csv [first_launch_id last_launch_id]"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr):
        continue
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    rows.append((int(r[ix["ID"]]), r[ix["Kernel Name"]], float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
sel = [r for r in rows if lo <= r[0] <= hi]
unit = sel[0][3] if sel else "?"
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, name, v, _ in sel:
    k = name.split("(")[0][:70]
    agg[k][0] += 1
    agg[k][1] += v * scale
tot = sum(v[1] for v in agg.values())
print(f"# launches {len(sel)} (ids {lo}..{hi}), total {tot / 1e3:.3f} ms (sum of per-launch durations, us = {unit})")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  avg {t / n:9.1f} us  {k}")
