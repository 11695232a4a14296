#!/usr/bin/env python
"""Summarise ncu output into the small text files kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
    python tools/ncu_summary.py traffic  gpurun_out/launches.csv "how"    > profiles/traffic.json
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep          > profiles/rNN_kernels.txt
    python tools/ncu_summary.py stalls   gpurun_out/prof.ncu-rep [kernel] > profiles/rNN_stalls.txt

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv`; `report` / `stalls` call
`ncu -i <rep> --page raw|source --csv` (ncu is in the image; no GPU needed to read a report).
"""
from __future__ import annotations

import collections
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def ncu_page(rep: str, page: str) -> list[list[str]]:
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        name = r[ki].split("(")[0].replace("void ", "").replace("ptd::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us serialised device time (cold cache, under ncu)")
    print(f"{'kernel':58s} {'n':>5s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} {v[0]:5d} {v[1]:10.1f} {v[1] / tot:7.3f} {v[1] / v[0]:8.1f}")


def traffic(path: str, source: str) -> None:
    """dram__bytes_read/write.sum per launch, averaged per kernel, as the JSON bench.py reads (profiles/traffic.json)."""
    import json

    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, mi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        except ValueError:
            continue
        if "_kernel<(bool)1" in r[ki] or "_kernel<1" in r[ki]:
            continue  # the counting variants of the traversal kernels (bench.py's counting pass): not what the timed region runs
        name = r[ki].split("(")[0].replace("void ", "").replace("ptd::<unnamed>::", "").replace("unnamed>::", "").split("<")[0]
        per[name][r[mi]] += v
        ids[name].add(r[ii])
    out = {}
    for name, m in sorted(per.items()):
        n = max(len(ids[name]), 1)
        rd, wr = m.get("dram__bytes_read.sum", 0.0) / n, m.get("dram__bytes_write.sum", 0.0) / n
        out[name] = {"launches": n, "dram_bytes_per_launch": rd + wr, "dram_read_bytes_per_launch": rd,
                     "dram_write_bytes_per_launch": wr, "avg_us_under_ncu": m.get("gpu__time_duration.sum", 0.0) / n}
        # issue-slot / lane / FP64-pipe figures when the capture carried them (tools/profile_launches.sh)
        inst, tinst = m.get("smsp__inst_executed.sum", 0.0), m.get("smsp__thread_inst_executed.sum", 0.0)
        if inst:
            out[name]["inst_per_launch"] = inst / n
            if tinst:
                out[name]["lanes_per_inst"] = tinst / inst
            f64 = m.get("smsp__inst_executed_pipe_fp64.sum", 0.0)
            if f64:
                out[name]["fp64_inst_per_launch"] = f64 / n
    out["_source"] = source
    print(json.dumps(out, indent=1))


def report(rep: str) -> None:
    rows = ncu_page(rep, "raw")
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== " + r[name_i].replace("ptd::<unnamed>::", "")[:110] + "  grid " + r[hdr.index("Grid Size")] + " block " + r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:82s} {r[i]:>16s} {units[i]}")


def stalls(rep: str, kernel: str | None) -> None:
    """Hottest source lines per kernel: warp-stall samples and executed instructions aggregated from the joined
    CUDA-C / SASS view (needs -lineinfo at compile time and --import-source on at capture time)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True, check=True).stdout
    fpath, func, hdr = None, None, None
    agg: dict = {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Function Name":
            func = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr is not None and len(r) == len(hdr) and r[0].isdigit():
            cols = {c: k for k, c in enumerate(hdr)}
            src_i = hdr.index("Source")
            try:
                samples = float(r[cols["# Samples"]] or 0)
                insts = float(r[cols["Instructions Executed"]] or 0)
                wait = float(r[cols.get("stall_wait", 0)] or 0) if "stall_wait" in cols else 0.0
                long_sb = float(r[cols["stall_long_sb"]] or 0) if "stall_long_sb" in cols else 0.0
            except ValueError:
                continue
            key = (func, os.path.basename(fpath or "?"), int(r[0]))
            a = agg.setdefault(key, [0.0, 0.0, 0.0, 0.0, r[src_i].strip()])
            a[0] += samples
            a[1] += insts
            a[2] += wait
            a[3] += long_sb
    funcs = sorted({k[0] for k in agg})
    for fn in funcs:
        if kernel and kernel not in fn:
            continue
        rows = [(k, v) for k, v in agg.items() if k[0] == fn]
        tot_s = sum(v[0] for _, v in rows) or 1.0
        tot_i = sum(v[1] for _, v in rows) or 1.0
        print(f"== {fn.replace('ptd::<unnamed>::', '')}: {tot_s:.0f} stall samples, {tot_i:.0f} warp instructions")
        print(f"   {'samples':>8s} {'insts':>7s} {'wait':>6s} {'longsb':>6s}  file:line  source")
        for k, v in sorted(rows, key=lambda kv: -kv[1][0])[:40]:
            print(f"   {v[0] / tot_s:8.3f} {v[1] / tot_i:7.3f} {v[2] / tot_s:6.3f} {v[3] / tot_s:6.3f}  {k[1]}:{k[2]}  {v[4][:110]}")


def main() -> None:
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[2])
    elif mode == "report":
        report(sys.argv[2])
    elif mode == "stalls":
        stalls(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
