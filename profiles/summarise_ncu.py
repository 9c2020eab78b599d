#!/usr/bin/env python
"""Trim an `ncu -i report.ncu-rep --page raw --csv` dump to the columns the roofline discussion uses.

usage: python profiles/summarise_ncu.py raw.csv summary.csv
(the capture itself: `ncu --set full --clock-control none --import-source on -k regex:<kernel> ...`, see
B200_PROFILING.md; the raw dump has ~2400 columns per launch)"""
import csv
import sys

COLUMNS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    idx = [names.index(c) for c in COLUMNS if c in names]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([names[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[head + 2:]:
            if len(r) == len(names):
                w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
