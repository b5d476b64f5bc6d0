#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into profiles/<name>.md + roofline_traffic.json.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_scan --rows 10000000 --dim 1024 --esize 4
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_prefix")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--esize", type=int, default=4)
    ap.add_argument("--traffic-json", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu summary of `{a.rep}`", "", "`ncu --set full --clock-control none --import-source on` (cold cache, serialised: compare shares, not absolutes).", ""]
    traffic = None
    for r in data:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for k in KEYS:
            if k in hdr:
                vals[k] = (r[hdr.index(k)], units[hdr.index(k)])
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines += ["", "warps stalled per issue-active cycle (top): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]), ""]
        try:
            rd = float(vals["dram__bytes_read.sum"][0]) * UNIT[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * UNIT[vals["dram__bytes_write.sum"][1]]
            dur = float(vals["gpu__time_duration.sum"][0]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[vals["gpu__time_duration.sum"][1]]
            algo = a.rows * a.dim * a.esize
            lines += [f"DRAM traffic per launch: {rd + wr:.4e} B (read {rd:.4e}, write {wr:.4e}); "
                      f"algorithmic bytes: {algo:.4e} B; ratio {((rd + wr) / algo) if algo else float('nan'):.4f}; "
                      f"under-profiler rate {(rd + wr) / dur / 1e9:.0f} GB/s", ""]
            if traffic is None:
                traffic = {"kernel": name, "rows": a.rows, "dim": a.dim, "esize": a.esize, "dram_bytes_per_launch": rd + wr,
                           "algorithmic_bytes_per_launch": algo, "source": a.rep}
        except Exception:
            pass
    open(a.out_prefix + ".md", "w").write("\n".join(lines))
    if a.traffic_json and traffic:
        json.dump(traffic, open(a.traffic_json, "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
