#!/usr/bin/env python
"""Summarise an `ncu --set full --page raw --csv` export of the GEMM launches of one train step
(tools/ncu_gemm.sh) into the JSON bench.py reads for `roofline.traffic`.
Usage: python tools/ncu_gemm_summary.py gpurun_out/<tag>/raw.csv out.json "<provenance text>" """
import csv
import json
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name):
        v = r[ix[name]].replace(",", "")
        return float(v) if v not in ("", "n/a") else 0.0

    kernels = []
    for r in rows[2:]:
        if len(r) < len(hdr) or "gemm_kernel" not in r[ix["Kernel Name"]]:
            continue
        m = re.search(r"<([^>]*)>", r[ix["Kernel Name"]])
        kernels.append({
            "template": "<BN,A_MN,B_MN,ACT,EPI_TMA,CTA2>=" + (m.group(1) if m else "?"),
            "us": col(r, "gpu__time_duration.sum"),
            "dram_read_MB": col(r, "dram__bytes_read.sum"),
            "dram_write_MB": col(r, "dram__bytes_write.sum"),
            "tensor_pipe_active_pct": col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lts_pct": col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        })
    units = dict(zip(hdr, rows[1]))
    scale = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}
    for k in kernels:          # ncu picks a unit per column: normalise to MB
        k["dram_read_MB"] *= scale.get(units["dram__bytes_read.sum"], 1.0)
        k["dram_write_MB"] *= scale.get(units["dram__bytes_write.sum"], 1.0)
        if units["gpu__time_duration.sum"].startswith("ms"):
            k["us"] *= 1e3
        elif units["gpu__time_duration.sum"].startswith("ns"):
            k["us"] *= 1e-3
    tot_us = sum(k["us"] for k in kernels)
    dram = sum(k["dram_read_MB"] + k["dram_write_MB"] for k in kernels)
    out = {"source": sys.argv[3] if len(sys.argv) > 3 else "",
           "launches": len(kernels), "gemm_us_per_step_under_ncu": tot_us, "dram_MB_per_step": dram,
           "traffic_bytes_per_launch": dram * 1e6 / max(len(kernels), 1),
           "time_weighted_tensor_pipe_active_pct": sum(k["us"] * k["tensor_pipe_active_pct"] for k in kernels) / max(tot_us, 1e-9),
           "kernels": kernels}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))


if __name__ == "__main__":
    main()
