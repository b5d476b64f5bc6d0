#!/usr/bin/env python
"""Where do the microseconds between `value` (queries resident, launches back to back) and `e2e` (psx_search: host query
in, host result out, one query at a time) go?  Times, on one GPU and the same index:

    A  psx_search                                   (H2D + scan + 2 D2H + sync)          per call
    B  psx_search_device + synchronize              (scan alone, one at a time)          per call
    C  psx_search_device x n, one synchronize       (back to back)                        per call
    D  an empty stream synchronize                  (host wake-up cost)

    gpurun -- python tools/e2e_overhead.py --rows 4000000
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from photo_search_engine_b200 import _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    rows, d, k = args.rows, bench.DIM, bench.TOPK
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    index = _native.NativeIndex(d, _native.METRIC_IP, _native.STORE_F32, 0)
    bench.build_corpus(torch, index, 0, rows, d, dev)
    queries = bench.make_queries(torch, 64, d, dev)
    qh = queries.cpu().numpy()
    scores = torch.empty((1, 128), device=dev, dtype=torch.float32)
    ids = torch.empty((1, 128), device=dev, dtype=torch.int64)
    keys = torch.empty((1, 128), device=dev, dtype=torch.int64)
    st = torch.cuda.current_stream().cuda_stream
    n = args.steps

    def dev_call(i):
        index.search_device(queries[i % 64: i % 64 + 1].data_ptr(), 1, k, scores.data_ptr(), ids.data_ptr(), keys.data_ptr(),
                            stream=st)

    out = {"rows": rows}
    for name in ("A", "B", "C", "A", "B", "C"):
        for i in range(10):
            index.search(qh[i], k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name == "A":
            for i in range(n):
                index.search(qh[i % 64], k)
        elif name == "B":
            for i in range(n):
                dev_call(i)
                torch.cuda.synchronize()
        else:
            for i in range(n):
                dev_call(i)
            torch.cuda.synchronize()
        out.setdefault(name, []).append((time.perf_counter() - t0) / n * 1e3)
    t0 = time.perf_counter()
    for i in range(1000):
        torch.cuda.synchronize()
    out["D_empty_sync_us"] = (time.perf_counter() - t0) * 1e3
    # host-side cost of the ctypes call and the numpy marshalling alone: k results of an EMPTY index
    empty = _native.NativeIndex(d, _native.METRIC_IP, _native.STORE_F32, 0)
    t0 = time.perf_counter()
    for i in range(1000):
        empty.search(qh[i % 64], k)
    out["python_call_on_empty_index_us"] = (time.perf_counter() - t0) * 1e3
    out["A_minus_B_us"] = (min(out["A"]) - min(out["B"])) * 1e3
    out["B_minus_C_us"] = (min(out["B"]) - min(out["C"])) * 1e3
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
