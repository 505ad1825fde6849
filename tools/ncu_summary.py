#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        d["kernel"] = name[:90]
        for i, h in enumerate(hdr):
            if h in WANT or "tensor" in h and "pct" in h or "stall" in h and "pct" in h and "warp_issue" in h:
                try:
                    d[h + " [" + units[i] + "]"] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[h] = vals[i]
        res.append(d)
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
