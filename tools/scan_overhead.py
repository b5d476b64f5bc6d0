#!/usr/bin/env python
"""Where does the scan's fixed cost go?  Sweeps the row count on one GPU, times back-to-back
launches with CUDA events, and dumps the per-CTA phase trace (psx_set_trace_device) of one launch.

    python tools/scan_overhead.py [--dim 1024] [--k 100] [--rows 10000,100000,1000000,...]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from photo_search_engine_b200 import _native  # noqa: E402


def fill(ix, rows, dim, gen):
    ix.reserve(rows)
    done = 0
    while done < rows:
        m = min(1 << 20, rows - done)
        blk = torch.randn((m, dim), generator=gen, device="cuda")
        blk /= blk.norm(dim=1, keepdim=True)
        ix.add_device(blk.data_ptr(), m)
        done += m


def timed(ix, qs, k, steps, flt=None):
    sc = torch.empty((1, k), device="cuda")
    ids = torch.empty((1, k), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()
    for i in range(5):
        ix.search_device(qs[i % len(qs)].data_ptr(), 1, k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=st.cuda_stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(steps):
        ix.search_device(qs[i % len(qs)].data_ptr(), 1, k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=st.cuda_stream)
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3  # us


def trace_once(ix, q, k, flt=None):
    buf = torch.zeros((148 * 8 * 8,), dtype=torch.int64, device="cuda")
    sc = torch.empty((1, k), device="cuda")
    ids = torch.empty((1, k), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream()
    ix.set_trace_device(buf.data_ptr())
    ix.search_device(q.data_ptr(), 1, k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ix.set_trace_device(0)
    both = buf.cpu().numpy().reshape(-1, 8)
    t, ph = both[:148], both[148:296]  # the kernel's stamps; phase 1 of a launch that compacts its own row list
    ph = ph[t[:, 0] != 0]
    t = t[t[:, 0] != 0]
    t0 = t[:, 0].min()
    rel = (t[:, :6] - t0) / 1e3
    last = int(np.argmax(t[:, 4]))
    pct = lambda v: [round(float(np.percentile(v, p)), 1) for p in (0, 50, 100)]
    return {
        "ctas": int(t.shape[0]),
        "entry_us[min,med,max]": pct(rel[:, 0]),
        "prologue_done_us": pct(rel[:, 1]),
        "stream_done_us": pct(rel[:, 2]),
        "published_us": pct(rel[:, 3]),
        "last_cta_merged_us": round(float(rel[last, 4]), 1),
        "last_cta_emitted_us": round(float(rel[last, 5]), 1),
        "compactions[min,med,max]": pct(t[:, 6]),
        **({"phase1_us[ticket,evaluated,reserved,written,all_done] med/max": [[round(float(np.median((ph[:, j][ph[:, j] != 0] - t0) / 1e3)), 1),
                                                                                 round(float(((ph[:, j][ph[:, j] != 0] - t0) / 1e3).max()), 1)]
                                                                                for j in range(5)]} if (ph[:, 4] != 0).any() else {}),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--rows", default="10000,100000,500000,1000000,1250000,2000000,4000000")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--filters", action="store_true", help="also time the EXIF predicate paths (fused vs row list)")
    a = ap.parse_args()
    gen = torch.Generator(device="cuda").manual_seed(3)
    qs = torch.randn((16, a.dim), generator=gen, device="cuda")
    qs /= qs.norm(dim=1, keepdim=True)
    for rows in [int(v) for v in a.rows.split(",")]:
        ix = _native.NativeIndex(a.dim, 0, _native.STORE_F32, 0)
        fill(ix, rows, a.dim, gen)
        out = {"rows": rows}
        for deal in (0, 1, 0, 1):  # A/B on the same box, interleaved
            ix.set_tunable("deal", deal)
            us = timed(ix, qs, a.k, a.steps)
            out.setdefault(f"deal={deal}", []).append([round(us, 1), round(rows * a.dim * 4 / us / 1e3)])
        for pdl in (0, 2, 0, 2):  # 2 = overlap also across calls (one query per call here, queries resident)
            ix.set_tunable("pdl", pdl)
            us = timed(ix, qs, a.k, a.steps)
            out.setdefault(f"pdl={pdl}", []).append([round(us, 1), round(rows * a.dim * 4 / us / 1e3)])
        out["trace"] = trace_once(ix, qs[0], a.k)
        print(json.dumps(out), flush=True)
        if a.filters:
            # dt words uniform in 1..1000: a [1, s] window passes s/1000 of the rows
            words = torch.randint(1, 1001, (rows,), generator=gen, device="cuda", dtype=torch.int64)
            ix.set_attrs_device(0, words.data_ptr(), rows)
            for sel in (800, 200, 70, 30, 5):
                flt = _native.PsxFilter(flags=_native.F_NEED_DT | _native.F_START | _native.F_END, start=1, end=sel)
                passing = int((words <= sel).sum())
                algo = passing * a.dim * 4 + rows * 8
                res = {"rows": rows, "pass_frac": sel / 1000}
                for mode in (1, 2, 3, 1, 2, 3):
                    ix.set_tunable("filter_mode", mode)
                    us = timed(ix, qs, a.k, a.steps, flt)
                    res.setdefault(f"filter_mode={mode}", []).append([round(us, 1), round(algo / us / 1e3)])
                ix.set_tunable("filter_mode", 3)
                res["trace_row_list_scan"] = trace_once(ix, qs[0], a.k, flt)   # phase times of the listed scan (no overlap while traced)
                print(json.dumps(res), flush=True)
        ix.close()
        del ix
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
