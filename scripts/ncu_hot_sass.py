"""Top stalled SASS instructions of one kernel in an .ncu-rep (needs a --set full capture).
Usage: python scripts/ncu_hot_sass.py rep.ncu-rep kernel-regex [top_n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x") or (r and r[0].isdigit())]
if not data:
    data = [r for r in rows[hi + 1:] if len(r) > ix["# Samples"]]


def num(r, k):
    try:
        return int(r[ix[k]].replace(",", "") or 0)
    except (ValueError, IndexError):
        return 0


tot = sum(num(r, "# Samples") for r in data)
print(f"# {rx}: {len(data)} SASS instructions, {tot} samples")
stall_keys = [k for k in hdr if k.startswith("stall_") and "(Not" not in k]
agg = {k: sum(num(r, k) for r in data) for k in stall_keys}
print("# stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:topn]
for i in order:
    r = data[i]
    n = num(r, "# Samples")
    st = sorted(((k[6:], num(r, k)) for k in stall_keys), key=lambda kv: -kv[1])[:2]
    print(f"{n:6d} {100 * n / max(tot, 1):5.1f}% #{i:5d} exec={num(r, 'Instructions Executed'):>9d} "
          f"{r[ix['Source']][:72]:72s} {st}")
