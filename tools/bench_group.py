#!/usr/bin/env python
"""ONE process, G GPUs: QPS of the drop-in boundary itself (psx_search / VectorStore.search on a psx_create_sharded
handle) over the headline workload, 10M x 1024 fp32 top-100, next to the one-process-per-GPU numbers of bench.py.

    gpurun --gpus 8 -- python tools/bench_group.py --out profiles/r2_group_n8.json

For every G in --gpus-list (default 1,2,4,8 up to the visible devices): build the corpus (same seeds as bench.py), check
bit-equality with a single-device index on a few queries, time single queries (host query in, host result out), the
drop-in VectorStore.search (list in, dicts out) and a 256-query batch.  Prints one JSON object.
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

import bench  # noqa: E402  (corpus seeds / generators)
from photo_search_engine_b200 import _native  # noqa: E402
from photo_search_engine_b200.vector_store import VectorStore  # noqa: E402


def build(index, rows, d, devices):
    index.reserve(rows)
    quota = -(-rows // len(devices))
    done = 0
    while done < rows:
        c, off = divmod(done, bench.CHUNK)
        take = min(bench.CHUNK - off, rows - done)
        dev = torch.device("cuda", devices[min(done // quota, len(devices) - 1)])
        with torch.cuda.device(dev):
            gen = torch.Generator(device=dev).manual_seed(bench.CORPUS_SEED + c)
            blk = torch.randn((bench.CHUNK, d), generator=gen, device=dev, dtype=torch.float32)[off: off + take]
            blk = (blk / blk.norm(dim=1, keepdim=True)).contiguous()
            index.add_device(blk.data_ptr(), take, stream=torch.cuda.current_stream(dev).cuda_stream)
            del blk
        done += take


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=bench.ROWS)
    ap.add_argument("--dim", type=int, default=bench.DIM)
    ap.add_argument("--k", type=int, default=bench.TOPK)
    ap.add_argument("--gpus-list", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    rows, d, k = args.rows, args.dim, args.k
    dev0 = torch.device("cuda", 0)
    queries = bench.make_queries(torch, bench.N_QUERIES, d, dev0).cpu().numpy()
    out = {"workload": bench.workload_name(rows, d, k), "visible_gpus": ndev, "results": {}}
    reference = None  # (scores, ids) of a few queries on one device
    for G in [int(x) for x in args.gpus_list.split(",")]:
        if G > ndev:
            continue
        devices = list(range(G))
        index = _native.NativeIndex(d, _native.METRIC_IP, _native.STORE_F32, devices if G > 1 else 0)
        t0 = time.perf_counter()
        build(index, rows, d, devices)
        build_s = time.perf_counter() - t0
        res = {"devices": devices, "build_s": build_s, "shard_rows": [r for _, r in index.shard_rows()]}
        if G == 1:
            index.set_tunable("batch_min", 0)
            D, I = index.search(queries[:8], k)
            index.set_tunable("batch_min", 4)
            reference = (D, I)
        elif reference is not None:
            got = [index.search(queries[i], k) for i in range(8)]
            res["bit_identical_to_one_device"] = bool(all(np.array_equal(g[1][0], reference[1][i]) and np.array_equal(g[0][0], reference[0][i])
                                                          for i, g in enumerate(got)))
        # single queries through the C ABI (psx_search: host query -> host scores / ids)
        for i in range(10):
            index.search(queries[i % 64], k)
        t0 = time.perf_counter()
        for i in range(args.steps):
            index.search(queries[i % 64], k)
        dt = time.perf_counter() - t0
        res["psx_search_qps"] = args.steps / dt
        res["psx_search_ms"] = dt / args.steps * 1e3
        res["scanned_GBps_aggregate"] = rows * d * 4 / (dt / args.steps) / 1e9
        # the drop-in class: list in, k dicts out
        store = VectorStore(None, "/tmp/_g.index", "/tmp/_g.json")
        store.dimension, store.index = d, index
        shared = {"photo_path": "synthetic"}
        store.metadata = [shared] * rows
        qs = [queries[i].tolist() for i in range(8)]
        for i in range(5):
            store.search(qs[i], k)
        t0 = time.perf_counter()
        n = max(20, args.steps // 4)
        for i in range(n):
            store.search(qs[i % 8], k)
        res["vector_store_search_qps"] = n / (time.perf_counter() - t0)
        store.index = None
        # a 256-query batch (tensor-core path per shard, key lists merged on the home device)
        gen = torch.Generator(device=dev0).manual_seed(bench.QUERY_SEED + 1)
        qb = torch.randn((256, d), generator=gen, device=dev0)
        qb = (qb / qb.norm(dim=1, keepdim=True)).cpu().numpy()
        for _ in range(2):
            index.search(qb, k)
        t0 = time.perf_counter()
        nb = 5
        for _ in range(nb):
            Db, Ib = index.search(qb, k)
        res["batch256_ms"] = (time.perf_counter() - t0) / nb * 1e3
        res["batch256_queries_per_s"] = 256 / res["batch256_ms"] * 1e3
        if G == 1:
            out["batch_reference"] = True
            batch_ref = (Db, Ib)
        else:
            res["batch_bit_identical_to_one_device"] = bool(np.array_equal(Ib, batch_ref[1]) and np.array_equal(Db, batch_ref[0]))
        if G > 1:
            res["group_stats_fused_keyed_timeouts"] = list(index.group_stats())
        res["batch_stats_queries_fallbacks"] = list(index.batch_stats())
        out["results"][str(G)] = res
        index.close()
        print(json.dumps({str(G): res}), file=sys.stderr, flush=True)
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
