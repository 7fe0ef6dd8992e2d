#!/usr/bin/env python
"""Long-format ncu CSV (--csv, one row per kernel x metric) -> one line per launch with the metrics side by side."""
import csv
import re
import sys
from collections import OrderedDict

SHORT = {
    "gpu__time_duration.sum": "us", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "fmaheavy%",
    "smsp__issue_active.avg.pct": "issue%", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
    "launch__registers_per_thread": "regs", "dram__bytes_read.sum": "dram_rd_MB", "dram__bytes_write.sum": "dram_wr_MB",
    "lts__t_sector_hit_rate.pct": "l2hit%", "smsp__inst_executed.sum": "Minst",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "st_long",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "st_math",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "st_wait",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "st_short",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "st_noinst",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "st_branch",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "st_lg",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum": "local_ld_MB", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum": "local_st_MB",
}


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[1:]:
        key = r[ix["ID"]]
        name = re.sub(r"^void vmsm_kernel<vmsm::|>\(.*$|\(.*$", "", r[ix["Kernel Name"]])
        d = launches.setdefault(key, {"kernel": name, "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        m, unit, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]].replace(",", "")
        if m not in SHORT:
            continue
        try:
            v = float(v)
        except ValueError:
            continue
        if SHORT[m] == "us":
            v = v / 1e3 if unit in ("ns", "nsecond") else v
        if SHORT[m].endswith("_MB"):
            v = v / 1e6 * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if SHORT[m] == "Minst":
            v /= 1e6
        d[SHORT[m]] = round(v, 3)
    cols = ["kernel", "grid", "block"] + [c for c in SHORT.values()]
    w = csv.writer(sys.stdout, lineterminator="\n")  # grid / block contain commas: quoted
    w.writerow(cols)
    for d in launches.values():
        w.writerow([d.get(c, "") for c in cols])


if __name__ == "__main__":
    main(sys.argv[1])
