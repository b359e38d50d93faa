#!/usr/bin/env python
"""Print selected metrics of every kernel in an `ncu --page raw --csv` export.
Usage: python tools/ncu_raw_pick.py raw.csv [substring ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum ", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__occupancy_limit", "launch__registers_per_thread ", "stalled_long_scoreboard_per", "stalled_barrier_per", "stalled_short_scoreboard_per",
           "stalled_wait_per", "stalled_math_pipe", "stalled_mio_throttle_per", "stalled_lg_throttle_per", "stalled_not_selected_per",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__waves"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2:] or DEFAULT
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[name_i][:100])
        for h, u, v in zip(hdr, units, r):
            if any(w in h + " " for w in want):
                print(f"   {h:90s} {u:12s} {v}")


if __name__ == "__main__":
    main()
