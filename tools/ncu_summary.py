#!/usr/bin/env python
"""Summarise an .ncu-rep -- or the `ncu -i ... --page raw --csv` export of one, which is what tools/gpu_profile.sh brings
back from the GPU box (reports are too large for gpurun_out/) -- into a small CSV that is committed under profiles/."""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main(rep, out):
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[hdr.index(m)]}]" for m in METRICS if m in hdr])
        for r in rows[2:]:
            w.writerow([r[hdr.index("Kernel Name")]] + [r[hdr.index(m)] for m in METRICS if m in hdr])
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
