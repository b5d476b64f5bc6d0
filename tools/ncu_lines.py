#!/usr/bin/env python
"""Join the SASS page of an .ncu-rep with nvdisasm line info: executed warp instructions and stall
samples per CUDA source line of one kernel (read here, without a GPU).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep photo_search_engine_b200/build/psx_scan_bf16_ip.o \
        '_ZN3psx16scan_topk_kernelI13__nv_bfloat16Li0ELi3ELb1ELi0EEEvNS_10ScanParamsE' --rows 12500000
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("obj")
    ap.add_argument("mangled")
    ap.add_argument("--rows", type=float, default=1.0, help="normalise counts by this many units")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=tmp, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    i0 = [i for i, l in enumerate(sass) if l.strip() == a.mangled + ":"][0]
    cur, seq = None, []
    for l in sass[i0 + 2:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            seq.append(cur)
        elif l.strip().startswith(".text.") or l.strip().startswith(".section"):
            break
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"] + [len(rows)]
    hdr = rows[his[a.launch]]
    data = [r for r in rows[his[a.launch] + 1: his[a.launch + 1]] if len(r) > 5 and r[0].startswith("0x")]
    ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    assert abs(len(data) - len(seq)) < 8, (len(data), len(seq))
    cnt, smp = collections.Counter(), collections.Counter()
    for k in range(min(len(data), len(seq))):
        cnt[seq[k]] += int(data[k][ie] or 0)
        smp[seq[k]] += int(data[k][ss] or 0)
    src = {}
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in os.listdir(os.path.join(here, "photo_search_engine_b200", "csrc")):
        src[f] = open(os.path.join(here, "photo_search_engine_b200", "csrc", f)).read().split("\n")
    tot, stot = sum(cnt.values()), sum(smp.values())
    print(f"total warp instructions {tot} = {tot / a.rows:.2f} per unit; stall samples {stot}")
    for key, c in cnt.most_common(a.top):
        f, ln = key if key else ("?", 0)
        text = src[f][ln - 1].strip()[:100] if f in src and 0 < ln <= len(src[f]) else ""
        print(f"{f}:{ln:<4d} {c / a.rows:8.2f}  {100.0 * smp[key] / max(stot, 1):5.1f}% stalls  {text}")


if __name__ == "__main__":
    main()
