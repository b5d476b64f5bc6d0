#!/usr/bin/env python
"""torchrun check (2+ GPUs): a query batch over row shards (tensor-core path per shard, one all-gather of the
keys, merge) equals the same batch scanned query by query over the whole corpus on one GPU, bit for bit.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_batch.py
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from photo_search_engine_b200 import _native  # noqa: E402
from photo_search_engine_b200.sharded import ShardedIndex, shard_bounds  # noqa: E402


def main():
    rank, world, local_rank = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    out = {}
    for n, d, nq, k, tier in ((400_000, 256, 64, 100, _native.STORE_F32), (300_000, 1024, 200, 50, _native.STORE_F32),
                              (300_000, 768, 130, 100, _native.STORE_BF16_MASTER), (100_000, 128, 9, 10, _native.STORE_F32)):
        g = torch.Generator(device=dev).manual_seed(7)  # same stream of numbers on every rank
        x = torch.randn((n, d), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        q = torch.randn((nq, d), generator=g, device=dev)
        q /= q.norm(dim=1, keepdim=True)
        q[0] = x[n - 5]
        words = torch.arange(n, device=dev, dtype=torch.int64) + 1
        bounds = shard_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        whole = _native.NativeIndex(d, 0, _native.STORE_F32, local_rank)
        whole.add_device(x.data_ptr(), n)
        whole.set_attrs_device(0, words.data_ptr(), n)
        whole.set_tunable("batch_min", 0)
        shard = _native.NativeIndex(d, 0, tier, local_rank)
        shard.add_device(x[lo:hi].contiguous().data_ptr(), hi - lo)
        shard.set_attrs_device(0, words[lo:hi].contiguous().data_ptr(), hi - lo)
        sh = ShardedIndex(shard, lo, bounds=bounds)
        for flt in (None, _native.PsxFilter(flags=_native.F_NEED_DT | _native.F_START | _native.F_END, start=n // 3, end=n - 3)):
            Dw, Iw = whole.search(q.cpu().numpy(), k, flt)
            S, I = sh.search(q.cpu().numpy(), k, flt)
            same = bool((I == Iw).all() and (S == Dw).all())
            out[f"n={n} d={d} nq={nq} k={k} tier={tier} filter={flt is not None}"] = same
        shard.close()
        whole.close()
    ok = torch.tensor([int(all(out.values()))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"all_ranks_ok": bool(ok.item()), "cases": out}, indent=1))
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
