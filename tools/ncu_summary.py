#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv  profiles/NAME_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/X.ncu-rep       profiles/NAME_ncu.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_alu.avg", "sm__inst_executed_pipe_fma.avg",
        "sm__inst_executed_pipe_xu", "sm__warps_active.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__throughput.avg", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum ", "sm__throughput.avg", "gpu__dram_throughput.avg", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum ", "smsp__issue_active.avg", "l1tex__throughput.avg", "sm__cycles_active.avg",
        "smsp__thread_inst_executed_per_inst_executed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "sm__sass_thread_inst_executed_op_integer_pred_on.sum ", "smsp__inst_executed_op_shared"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, agg, order = None, defaultdict(lambda: [0, 0.0]), []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        k = d["Kernel Name"].split("(")[0][-70:]
        agg[k][0] += 1
        agg[k][1] += ms
        order.append((d["ID"], k, ms))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; %d launches, %.3f ms total\n" % (src, len(order), tot))
        f.write("%-72s %5s %12s %7s\n" % ("kernel", "n", "ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %5d %12.3f %6.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
        f.write("\n# per launch, in order\n")
        for i, k, ms in order:
            f.write("%4s %-72s %10.3f ms\n" % (i, k, ms))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        out.append("## %s  grid %s block %s" % (d.get("Kernel Name", "?")[:100], d.get("Grid Size", "?"), d.get("Block Size", "?")))
        for h, u, v in zip(hdr, units, vals):
            if any(k in h + " " for k in KEYS) and ".max" not in h.replace("elapsed.max", "") and ".min" not in h and ".sum.p" not in h:
                out.append("%-95s %-14s %s" % (h, u, v))
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src_csv)))
    hdr = next((r for r in rows if "Source" in r and "# Samples" in r), None)
    if hdr:
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[ix["# Samples"]].isdigit()]
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        agg = sorted(((sum(int(r[ix[s]] or 0) for r in data), s) for s in stalls), reverse=True)
        out.append("\n## warp-stall sampling, all warps: %d samples" % tot)
        out.append("  ".join("%s %.1f%%" % (s, 100.0 * n / tot) for n, s in agg[:8]))
        out.append("\n## hottest SASS instructions (samples, executed, instruction, top stall)")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
            st = max(((int(r[ix[s]] or 0), s) for s in stalls))
            out.append("%8s %12s  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:70], st[1]))
    open(dst, "w").write("# ncu --set full --clock-control none --import-source on ; source: %s\n" % src + "\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
