#!/usr/bin/env python
"""Sweep the scan kernel's launch geometry on a resident synthetic corpus (GPU box only).

    python tools/tune_scan.py --rows 2000000 --dim 1024 --k 100 [--store bf16] [--filter]
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from photo_search_engine_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--store", default="fp32")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warps", default="4,8,16")
    ap.add_argument("--stages", default="2,3,4,6")
    ap.add_argument("--ctas", default="1,2")
    ap.add_argument("--filter", action="store_true")
    a = ap.parse_args()
    dt = _native.STORE_BF16 if a.store == "bf16" else _native.STORE_F32
    esize = 2 if a.store == "bf16" else 4
    ix = _native.NativeIndex(a.dim, 0, dt, 0)
    ix.reserve(a.rows)
    g = torch.Generator(device="cuda").manual_seed(1)
    done = 0
    while done < a.rows:
        m = min(1 << 20, a.rows - done)
        blk = torch.randn((m, a.dim), generator=g, device="cuda")
        blk /= blk.norm(dim=1, keepdim=True)
        ix.add_device(blk.data_ptr(), m)
        done += m
    q = torch.randn((16, a.dim), generator=g, device="cuda")
    q /= q.norm(dim=1, keepdim=True)
    flt = None
    if a.filter:
        words = torch.arange(a.rows, device="cuda", dtype=torch.int64) + 1
        ix.set_attrs_device(0, words.data_ptr(), a.rows)
        flt = _native.PsxFilter(flags=_native.F_NEED_DT | _native.F_START, start=1)
    sc = torch.empty((1, a.k), device="cuda")
    ids = torch.empty((1, a.k), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream()
    results = []
    for w, s, c in itertools.product(*[[int(v) for v in t.split(",")] for t in (a.warps, a.stages, a.ctas)]):
        ix.set_tunable("warps", w)
        ix.set_tunable("stages", s)
        ix.set_tunable("ctas_per_sm", c)
        try:
            for i in range(3):
                ix.search_device(q[i:i + 1].data_ptr(), 1, a.k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(a.steps):
                ix.search_device(q[i % 16:i % 16 + 1].data_ptr(), 1, a.k, sc.data_ptr(), ids.data_ptr(), 0, flt=flt, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            gbs = a.rows * a.dim * esize / ms / 1e6
            results.append(dict(warps=w, stages=s, ctas=c, ms=ms, GBps=gbs))
            print(f"warps={w:2d} stages={s} ctas/sm={c}  {ms:8.4f} ms  {gbs:8.1f} GB/s", flush=True)
        except Exception as exc:
            print(f"warps={w} stages={s} ctas/sm={c}  failed: {exc}", flush=True)
    best = max(results, key=lambda r: r["GBps"])
    print("best", json.dumps(best))


if __name__ == "__main__":
    main()
