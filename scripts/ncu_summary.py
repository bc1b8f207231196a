"""Compact text summary of an .ncu-rep (one line block per captured launch) for profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.txt"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none summary of {path}")
    for r in rows[2:]:
        print(f"\n== {r[ix['Kernel Name']][:100]}")
        for k, name in KEYS:
            if k in ix:
                print(f"   {name:26s} {r[ix[k]]:>16s} {units[ix[k]]}")


if __name__ == "__main__":
    main(sys.argv[1])
