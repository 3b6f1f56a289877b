#!/usr/bin/env python
"""Key metrics and stall reasons per kernel from an .ncu-rep (ncu --set full), via `ncu -i REP --page raw --csv`.
usage: ncu_extract.py file.ncu-rep > summary.txt"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shfl... ".strip()]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("--- %s" % r[col["Kernel Name"]][:100])
    for k in KEYS:
        if k in col:
            print("  %-72s %s %s" % (k, r[col[k]], units[col[k]]))
    st = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled_"):
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  top stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
