#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the text summaries committed under profiles/.

  tools/ncu_summaries.py launches <launches.csv> <out.txt>     per-kernel time / instructions / issue rate of ONE E-step
                                                               (ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum)
  tools/ncu_summaries.py full <report.ncu-rep> <out.txt>       key metrics + stall reasons + hottest instructions per kernel
                                                               (ncu --set full --import-source on)"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    iK, iM, iV, iI, iU = [h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit")]
    L = {}
    for r in rows[hi + 1:]:
        if len(r) <= iV:
            continue
        d = L.setdefault(int(r[iI]), {"name": re.sub(r"\(.*", "", r[iK]).replace("void ", "")})
        v = float(r[iV].replace(",", ""))
        if r[iM] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
        d[r[iM]] = v
    seq = [(L[i]["name"], L[i].get("gpu__time_duration.sum", 0.0), L[i].get("smsp__inst_executed.sum", 0.0)) for i in sorted(L)]
    starts = [k for k, s in enumerate(seq) if s[0].startswith("k_backward_warm")]
    with open(out, "w") as fo:
        fo.write("# per-launch device times are serialised and cold-cache under ncu: compare SHARES, not absolutes\n")
        fo.write("# (in production k_backward_warm and the two 'predicted' k_transfer launches run on a side stream, concurrently with k_forward)\n")
        for e, (a, b) in enumerate(zip(starts, starts[1:] + [len(seq)])):
            agg = collections.OrderedDict()
            for n, t, ins in seq[a:b]:
                d = agg.setdefault(n, [0, 0.0, 0.0])
                d[0] += 1; d[1] += t; d[2] += ins
            tot = sum(v[1] for v in agg.values()); ti = sum(v[2] for v in agg.values())
            fo.write("\n== E-step %d: %d launches, %.3f ms serialised, %.2f G warp instructions\n" % (e, b - a, tot, ti / 1e9))
            fo.write("%-36s %3s %10s %7s %9s %6s\n" % ("kernel", "n", "total ms", "share", "G instr", "IPC*"))
            for n, v in agg.items():
                ipc = v[2] / (v[1] * 1e-3 * 592 * 1.965e9) if v[1] > 0 else 0.0
                fo.write("%-36s %3d %10.4f %6.1f%% %9.4f %6.2f\n" % (n[:36], v[0], v[1], 100 * v[1] / tot, v[2] / 1e9, ipc))
        fo.write("\n* IPC per SM sub-partition = smsp__inst_executed.sum / (time x 592 SMSPs x 1.965 GHz)\n")


WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}; blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    with open(out, "w") as fo:
        for ki, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            fo.write("--- %s\n" % d["Kernel Name"][:150])
            for k in WANT:
                if k in d:
                    fo.write("  %-70s %s %s\n" % (k, d[k], u.get(k, "")))
            st = {k: v for k, v in d.items() if "smsp__average_warps_issue_stalled" in k and "_per_issue_active" in k and "not_issued" not in k}
            top = sorted(((float(v) if v not in ("", "n/a") else 0.0, k) for k, v in st.items()), reverse=True)[:8]
            fo.write("  top stalls (warps per issue): " + ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in top) + "\n")
            for b in blocks[ki:ki + 1]:   # (the source page lists the kernels in the same order)
                h = b["hdr"]; ia, isrc, ins, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
                tot = sum(int(x[ins] or 0) for x in b["rows"]) or 1
                fo.write("  hottest instructions (stall samples; wait / short_scoreboard / long_scoreboard / math_pipe):\n")
                for x in sorted(b["rows"], key=lambda x: -int(x[ins] or 0))[:12]:
                    fo.write("    %-58s %5.1f%%  exec %9s  wait %5s short %5s long %5s math %5s\n" % (x[isrc][:58], 100 * int(x[ins]) / tot, x[iex], x[h.index("stall_wait")],
                             x[h.index("stall_short_sb")], x[h.index("stall_long_sb")], x[h.index("stall_math")]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
