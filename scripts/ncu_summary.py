#!/usr/bin/env python
"""Summarise ncu artefacts (read here, no GPU needed):
   ncu_summary.py launches <launches.csv>            -> per-kernel count / avg / share
   ncu_summary.py raw <report.ncu-rep>               -> key metrics per captured launch
   ncu_summary.py stalls <report.ncu-rep> [top]      -> stall reasons + hottest SASS lines"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'avg_ns':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):5d} {sum(v)/len(v):10.0f} {sum(v)/tot:7.3f}")


def raw(path):
    rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")


def stalls(path, top=14):
    rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv"]))))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ci = {n: i for i, n in enumerate(hdr)}
    st = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
    agg = {s: sum(int(r[ci[s]] or 0) for r in data) for s in st}
    print("total samples", tot)
    print({k: f"{v} ({100*v/max(tot,1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:top]:
        s = {k: int(r[ci[k]] or 0) for k in st if int(r[ci[k]] or 0) > 0}
        print(f"{r[ci['# Samples']]:>6s}  {r[ci['Source']][:78]:78s} {dict(sorted(s.items(), key=lambda kv: -kv[1])[:2])}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "raw":
        raw(sys.argv[2])
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)
