#!/usr/bin/env python
"""Summarise an ncu report (`ncu --set full`) into the few numbers profiles/ keeps per round.

    python profiles/ncu_summary.py gpurun_out/prof_pixels.ncu-rep > profiles/rNN_<tag>_ncu_summary.txt

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and prints, per
captured launch: duration, DRAM bytes read/written, issue utilisation, occupancy, the
per-issue stall reasons and the pipe mix.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__inst_executed_pipe_fp64.sum", "pipe fp64"),
    ("sm__inst_executed_pipe_fma.sum", "pipe fma"),
    ("sm__inst_executed_pipe_alu.sum", "pipe alu"),
    ("sm__inst_executed_pipe_lsu.sum", "pipe lsu"),
    ("sm__inst_executed_pipe_xu.sum", "pipe xu"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launches (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']].split('(')[0]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"  {label:34s} {r[idx[k]]} {units[idx[k]]}")


if __name__ == "__main__":
    main()
