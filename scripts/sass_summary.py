"""SASS opcode summary of every kernel in libpainter_b200.so (cuobjdump): the tcgen05 / TMEM / TMA mnemonics that show
what a kernel really runs on, plus registers and static resource usage.
Usage: python scripts/sass_summary.py > profiles/rNN_sass_opcode_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "painter_b200", "libpainter_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS",
        "MUFU.EX2", "MUFU.TANH", "FFMA2", "FADD2", "FMUL2", "HMMA", "RED", "ATOM", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        c["_total"] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
        for k in KEYS:
            if k == "UTCHMMA.2CTA":
                continue
            if op == k or op.startswith(k + ".") or (k in ("RED", "ATOM") and op.startswith(k)):
                c[k] += 1
    names = demangle(list(counts))
    print(f"# SASS opcode summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass / -res-usage, sm_100a)")
    print("# per kernel: static instruction counts of the tcgen05 (UTCHMMA / UTCQMMA), tensor-memory (LDTM / STTM), TMA (UTMALDG /")
    print("# UTMASTG), mbarrier (SYNCS, UTCBAR) and packed-fp32 / MUFU mnemonics; registers per thread, stack bytes (spills)")
    for mangled, c in counts.items():
        if c["_total"] == 0:
            continue
        name = re.sub(r"\(.*", "", names.get(mangled, mangled))
        r = regs.get(mangled, ("?", "?", "?"))
        parts = [f"{k}={c[k]}" for k in KEYS if c[k]]
        print(f"{name[:64]:64s} instr={c['_total']:6d} regs={r[0]} stack={r[1]}  " + " ".join(parts))


if __name__ == "__main__":
    main()
