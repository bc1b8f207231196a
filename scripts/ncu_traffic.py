"""DRAM traffic per launch of the block-level GEMMs from an `ncu --set full` capture of `scripts/prof_kernels.py gemm`
(the LAST eight gemm2 launches of the report = the second, warm pass).  Writes the JSON `bench.py` reads for
`roofline.traffic`.
Usage: python scripts/ncu_traffic.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_gemm_traffic.json"""
import csv
import json
import subprocess
import sys

M, C = 12544, 1024
LABELS = [   # order of scripts/prof_kernels.py; algorithmic bytes = operands read once + result written once
    ("qkv fwd (bias->bf16)", 2 * M * C + 2 * 3 * C * C + 2 * M * 3 * C),
    ("proj fwd (residual)", 2 * M * C + 2 * C * C + 4 * M * C + 4 * M * C),
    ("fc1 fwd (GELU)", 2 * M * C + 2 * 4 * C * C + 2 * 2 * M * 4 * C),
    ("fc2 fwd (residual)", 2 * M * 4 * C + 2 * 4 * C * C + 4 * M * C + 4 * M * C),
    ("fc2 dgrad (gelu')", 2 * M * C + 2 * 4 * C * C + 2 * 2 * M * 4 * C),
    ("fc1 dgrad (fp32)", 2 * M * 4 * C + 2 * 4 * C * C + 4 * M * C),
    ("fc1 wgrad (stream-K)", 2 * M * 4 * C + 2 * M * C + 4 * 4 * C * C),
    ("qkv wgrad (stream-K)", 2 * M * 3 * C + 2 * M * C + 4 * 3 * C * C),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        return float(r[ix[k]].replace(",", ""))

    g = [r for r in rows[2:] if "gemm2_bf16_kernel" in r[ix["Kernel Name"]]][-len(LABELS):]
    units = rows[1]
    dscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}[
        units[ix["gpu__time_duration.sum"]]]
    bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[ix["dram__bytes_read.sum"]]]
    launches = []
    for (label, alg), r in zip(LABELS, g):
        launches.append({"launch": label, "kernel": r[ix["Kernel Name"]][:32],
                         "dram_bytes": (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) * bscale,
                         "algorithmic_bytes": float(alg), "duration_us": num(r, "gpu__time_duration.sum") * dscale})
    print(json.dumps({
        "source": "ncu --set full --clock-control none, scripts/prof_kernels.py gemm (one launch of each block-level "
                  "GEMM at B=8, M=12544, warm pass), dram__bytes_read.sum + dram__bytes_write.sum",
        "launches": launches,
        "avg_dram_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / max(1, len(launches))}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
