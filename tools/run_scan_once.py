#!/usr/bin/env python
"""A few single-query scans on a synthetic resident corpus: the command profiled under ncu.

    python tools/run_scan_once.py --rows 12500000 --dim 768 --store bf16 [--steps 6] [--pass-frac 0.07]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from photo_search_engine_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--store", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--pass-frac", type=float, default=0.0, help="> 0: EXIF window predicate passing this fraction of the rows")
    ap.add_argument("--tunable", action="append", default=[], help="key=value, repeatable")
    ap.add_argument("--ballast-gb", type=float, default=0.0, help="allocate this much HBM first (placement experiment)")
    ap.add_argument("--queries", type=int, default=8)
    a = ap.parse_args()
    ballast = torch.empty(int(a.ballast_gb * (1 << 30)), dtype=torch.uint8, device="cuda") if a.ballast_gb > 0 else None
    dt = _native.STORE_BF16 if a.store == "bf16" else _native.STORE_F32
    esize = 2 if a.store == "bf16" else 4
    ix = _native.NativeIndex(a.dim, 0, dt, 0)
    for kv in a.tunable:
        key, val = kv.split("=")
        ix.set_tunable(key, int(val))
    ix.reserve(a.rows)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < a.rows:
        m = min(1 << 20, a.rows - done)
        blk = torch.randn((m, a.dim), generator=g, device="cuda")
        blk /= blk.norm(dim=1, keepdim=True)
        ix.add_device(blk.data_ptr(), m)
        done += m
    q = torch.randn((a.queries, a.dim), generator=g, device="cuda")
    q /= q.norm(dim=1, keepdim=True)
    flt = None
    passing = a.rows
    if a.pass_frac > 0:
        words = torch.randint(1, 1001, (a.rows,), generator=g, device="cuda", dtype=torch.int64)
        ix.set_attrs_device(0, words.data_ptr(), a.rows)
        end = max(1, int(round(a.pass_frac * 1000)))
        flt = _native.PsxFilter(flags=_native.F_NEED_DT | _native.F_START | _native.F_END, start=1, end=end)
        passing = int((words <= end).sum())
    sc = torch.empty((1, a.k), device="cuda")
    ids = torch.empty((1, a.k), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        ix.search_device(q[i % a.queries].data_ptr(), 1, a.k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0.record(st)
    for i in range(a.steps):
        ix.search_device(q[i % a.queries].data_ptr(), 1, a.k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    algo = passing * a.dim * esize + (a.rows * 8 if flt is not None else 0)
    print(json.dumps({"rows": a.rows, "dim": a.dim, "store": a.store, "pass_rows": passing, "ms_per_query": ms,
                      "algorithmic_GBps": algo / ms / 1e6}))
    ix.close()


if __name__ == "__main__":
    main()
