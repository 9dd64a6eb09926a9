#!/usr/bin/env python
"""Condenses an `ncu --set full` report (read here, without a GPU: `ncu -i X.ncu-rep --page raw --csv`) into the table
kept under profiles/: per launch — duration, grid, registers, achieved occupancy, issue-slot utilisation, active lanes
per instruction (warp execution efficiency), DRAM bytes, and the top warp-stall reasons (cycles per issued instruction).

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
"""
import csv
import io
import re
import subprocess
import sys

COLS = [
    ("kernel", "Kernel Name"), ("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"), ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("lanes/32", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("dramRdMB", "dram__bytes_read.sum"), ("dramWrMB", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("Minst", "smsp__inst_executed.sum"),
    ("st_noinst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("st_longsb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("st_shortsb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("localLdSect", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
    ("l2hit%", "lts__t_sector_hit_rate.pct"),
]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main() -> None:
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(" ".join(f"{n:>11s}" for n, _ in COLS))
    for r in rows[2:]:
        out = []
        for name, col in COLS:
            if col not in idx:
                out.append(f"{'NA':>11s}")
                continue
            v = r[idx[col]]
            if name == "kernel":
                v = re.sub(r".*::", "", re.sub(r"\(.*", "", v))[:11]
            else:
                try:
                    f = float(v.replace(",", "")) * SCALE.get(units[idx[col]], 1.0)
                    if name == "Minst":
                        f /= 1e6
                    v = f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}"
                except ValueError:
                    pass
            out.append(f"{v:>11s}")
        print(" ".join(out))


if __name__ == "__main__":
    main()
