#!/usr/bin/env python
"""Time the tensor-core batched path (BASELINE config 3: 1M x 1024 fp32, nq=256, top-100)."""
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
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--clustered", action="store_true")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--trials", type=int, default=0, help="extra batches of fresh random queries: count unproven queries")
    ap.add_argument("--pair", action="store_true", help="cta_group::2 kernel for nq > 128")
    ap.add_argument("--store", default="fp32", choices=["fp32", "mixed"], help="mixed = bf16 rows + fp32 master (bf16 GEMM)")
    ap.add_argument("--tunable", action="append", default=[], help="key=value, repeatable")
    a = ap.parse_args()
    ix = _native.NativeIndex(a.dim, 0, _native.STORE_BF16_MASTER if a.store == "mixed" else _native.STORE_F32, 0)
    for kv in a.tunable:
        key, val = kv.split("=")
        ix.set_tunable(key, int(val))
    ix.reserve(a.rows)
    if a.pair:
        ix.set_tunable("batch_pair", 1)
    g = torch.Generator(device="cuda").manual_seed(a.seed)
    done = 0
    cent = torch.randn((4096, a.dim), generator=g, device="cuda")
    while done < a.rows:
        m = min(1 << 20, a.rows - done)
        blk = torch.randn((m, a.dim), generator=g, device="cuda")
        if a.clustered:
            blk = cent[torch.randint(0, 4096, (m,), generator=g, device="cuda")] + 0.35 * blk
        blk /= blk.norm(dim=1, keepdim=True)
        ix.add_device(blk.data_ptr(), m)
        done += m
    q = torch.randn((a.nq, a.dim), generator=g, device="cuda")
    q /= q.norm(dim=1, keepdim=True)
    sc = torch.empty((a.nq, a.k), device="cuda")
    ids = torch.empty((a.nq, a.k), dtype=torch.int64, device="cuda")
    flags = torch.zeros((a.nq,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def run():
        ix.search_batch_device(q.data_ptr(), a.nq, a.k, sc.data_ptr(), ids.data_ptr(), flags.data_ptr(), stream=stream.cuda_stream)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    flops = 2.0 * a.rows * a.dim * a.nq
    print(json.dumps({"rows": a.rows, "dim": a.dim, "nq": a.nq, "k": a.k, "store": a.store, "ms_per_batch": ms, "queries_per_s": a.nq / ms * 1e3,
                      "TFLOPs": flops / ms / 1e9, "GBps_corpus": a.rows * a.dim * 4 / ms / 1e6,
                      "unproven": int((flags != 0).sum()),
                      "unproven_detail": [(int(i), int(flags[i])) for i in flags.nonzero().flatten().tolist()][:8]}))
    if a.trials:
        bad = {}
        for t in range(a.trials):
            q2 = torch.randn((a.nq, a.dim), generator=g, device="cuda")
            q2 /= q2.norm(dim=1, keepdim=True)
            ix.search_batch_device(q2.data_ptr(), a.nq, a.k, sc.data_ptr(), ids.data_ptr(), flags.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            for i in flags.nonzero().flatten().tolist():
                bad[int(flags[i])] = bad.get(int(flags[i]), 0) + 1
        print(json.dumps({"trials": a.trials, "queries": a.trials * a.nq, "unproven_by_code": bad}))
        run()
        torch.cuda.synchronize()
    # correctness spot check vs the streaming scan
    ix.set_tunable("batch_min", 0)
    Ds, Is = ix.search(q[:8].cpu().numpy(), a.k)
    ok = flags[:8].cpu().numpy() == 0
    print("spot check ids equal:", bool((ids[:8].cpu().numpy()[ok] == Is[ok]).all()), "scores equal:", bool((sc[:8].cpu().numpy()[ok] == Ds[ok]).all()))


if __name__ == "__main__":
    main()
