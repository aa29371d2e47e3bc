#!/usr/bin/env python3
"""Key metrics of the first kernel of an .ncu-rep as JSON (for profiles/).  python tools/ncu_summary.py REP"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__cycles_active.avg"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, val = rows[0], rows[1], rows[2]
    out = {"source": rep}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            v = val[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            out[k + (" [%s]" % units[i] if units[i] else "")] = v
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            st[h[len(STALL):-len("_per_issue_active.ratio")]] = round(float(val[i]), 3)
    out["warps_stalled_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
