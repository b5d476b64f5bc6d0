"""Body of tests/test_gpu_torchrun.py.  Launched with ``python -m torch.distributed.run --nproc-per-node W`` on a box
with W >= 2 GPUs (one rank per GPU, NCCL), or as a plain script on a one-GPU box, where the W ranks are emulated on
device 0 through the same C-ABI entry points the ranks use (``psx_search_exchange_device`` with the phases split,
``psx_search_batch_device`` keys + ``psx_merge_keys_device``).  Either way it checks, against ONE index holding the whole
corpus (which test_gpu_parity.py checks against the oracle) and against the CPU oracle itself:

  (a) ShardedIndex.search, single queries          -- bit-identical ids and scores on every rank
  (b) a 64-query batch (tensor-core path per shard) -- bit-identical
  (c) search_by_id (image -> image)                 -- equals the whole-index search with the row itself dropped
  (d) config-5 miniature: bf16 row shards, recall@100 against the fp32 CPU oracle on >= 64 queries

Prints ``RESULT {json}`` on rank 0.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import c_oracle  # noqa: E402  (the checker)
from photo_search_engine_b200 import _native as N  # noqa: E402
from photo_search_engine_b200.sharded import ShardedIndex, shard_bounds  # noqa: E402


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def corpus(n, d, seed):
    rng = np.random.default_rng(seed)
    x = unit(rng, n, d)
    x[n - 7] = x[11]  # an exact tie whose two rows live on different shards
    return x, rng


def recall(got, want):
    return float(np.mean([len(set(g.tolist()) & set(w.tolist())) / len(w) for g, w in zip(got, want)]))


def real_ranks():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"mode": f"{world} ranks, one per GPU (NCCL)", "world": world}
    n, d, k = 290_000, 128, 100
    x, rng = corpus(n, d, 7)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    whole = N.NativeIndex(d, 0, 0, local)
    whole.add(x)
    whole.set_tunable("batch_min", 0)
    local_ix = N.NativeIndex(d, 0, 0, local)
    local_ix.add(x[lo:hi])
    sh = ShardedIndex(local_ix, lo, bounds=bounds)
    out["exchange"] = sh.exchange
    q = unit(rng, 64, d)
    q[0] = x[11]
    Dw, Iw = whole.search(q, k)
    ok = True
    for i in range(6):  # (a)
        D, I = sh.search(q[i], k)
        ok &= bool(np.array_equal(I[0], Iw[i]) and np.array_equal(D[0], Dw[i]))
    out["single_equal"] = ok
    out["tie_order"] = [int(Iw[0, 0]), int(Iw[0, 1])]
    D, I = sh.search(q, k)  # (b)
    out["batch_equal"] = bool(np.array_equal(I, Iw) and np.array_equal(D, Dw))
    out["batch_took_tensor_path"] = local_ix.batch_stats()[0] >= 64
    ok = True
    for gid in (5, lo, hi - 1, n - 7, int(rng.integers(0, n))):  # (c): every rank must pass the same ids
        gid = int(np.clip(gid, 0, n - 1))
        t = torch.tensor([gid], device=dev)
        dist.broadcast(t, src=0)
        gid = int(t.item())
        s, ids = sh.search_by_id(gid, k)
        De, Ie = whole.search(x[gid], k + 1)
        keep = Ie[0] != gid
        ok &= bool(np.array_equal(ids.cpu().numpy(), Ie[0][keep][:k]) and np.array_equal(s.cpu().numpy(), De[0][keep][:k]))
    out["by_id_equal"] = ok
    whole.close()
    # (d) config-5 miniature
    n5, d5, nq5 = 300_000, 768, 64
    x5, rng5 = corpus(n5, d5, 9)
    b5 = shard_bounds(n5, world)
    l5 = N.NativeIndex(d5, 0, N.STORE_BF16, local)
    l5.add(x5[b5[rank]: b5[rank + 1]])
    s5 = ShardedIndex(l5, b5[rank], bounds=b5)
    ids5 = rng5.integers(0, n5, nq5)
    got = []
    for gid in ids5.tolist():
        _, ids = s5.search_by_id(int(gid), k)
        got.append(ids.cpu().numpy())
    if rank == 0:
        c_oracle.set_threads(c_oracle.host_cores())
        _, Io = c_oracle.search(x5, x5[ids5], k + 1, nthreads=c_oracle.host_cores())
        want = [row[row != gid][:k] for row, gid in zip(Io, ids5)]
        out["config5_recall_at_100_vs_fp32_oracle"] = recall(got, want)
        out["config5_queries"] = nq5
        print("RESULT " + json.dumps(out), flush=True)
    local_ix.close()
    l5.close()
    dist.barrier()
    dist.destroy_process_group()


def emulated(world=3):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    out = {"mode": f"{world} ranks emulated on one GPU", "world": world}
    n, d, k = 290_000, 128, 100
    x, rng = corpus(n, d, 7)
    bounds = shard_bounds(n, world)
    whole = N.NativeIndex(d)
    whole.add(x)
    whole.set_tunable("batch_min", 0)
    shards = []
    for r in range(world):
        ix = N.NativeIndex(d)
        ix.add(x[bounds[r]: bounds[r + 1]])
        shards.append(ix)
    q = unit(rng, 64, d)
    q[0] = x[11]
    Dw, Iw = whole.search(q, k)
    qd = torch.from_numpy(q).to(dev)
    kp = N.kpad(k)
    # (a) fused exchange: every rank's scan publishes into every rank's buffer, then every rank waits + merges
    nwords = (N.exchange_bytes() + 7) // 8
    bufs = [torch.zeros(nwords, dtype=torch.int64, device=dev) for _ in range(world)]
    bases = np.array([b.data_ptr() for b in bufs], dtype=np.uint64)
    ok = True
    for seq in range(1, 7):
        qptr = qd.data_ptr() + (seq - 1) * d * 4
        for r in range(world):
            shards[r].search_exchange_device(qptr, k, r, world, bases, seq, 0, 0, id_base=bounds[r], stream=st, phases=1)
        for r in range(world):
            sc = torch.empty((1, k), dtype=torch.float32, device=dev)
            ids = torch.empty((1, k), dtype=torch.int64, device=dev)
            shards[r].search_exchange_device(0, k, r, world, bases, seq, sc.data_ptr(), ids.data_ptr(), stream=st, phases=2)
            torch.cuda.synchronize()
            ok &= bool(np.array_equal(ids.cpu().numpy()[0], Iw[seq - 1]) and np.array_equal(sc.cpu().numpy()[0], Dw[seq - 1]))
            ok &= shards[r].exchange_status() == 0
    out["single_equal"] = ok
    out["exchange"] = "p2p (emulated)"
    out["tie_order"] = [int(Iw[0, 0]), int(Iw[0, 1])]
    # (b) the batch: per-shard tensor-core keys, "all-gather" = the lists side by side, one merge CTA per query
    nq = 64
    lists = torch.zeros((nq, world, kp), dtype=torch.int64, device=dev)
    took = True
    for r in range(world):
        sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
        flags = torch.zeros((nq,), dtype=torch.int32, device=dev)
        mine = torch.zeros((nq, kp), dtype=torch.int64, device=dev)
        before = shards[r].batch_stats()[0]
        shards[r].search_batch_device(qd.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), flags.data_ptr(), mine.data_ptr(),
                                      id_base=bounds[r], stream=st)
        torch.cuda.synchronize()
        took &= shards[r].batch_stats()[0] - before == nq
        for qi in flags.cpu().numpy().nonzero()[0].tolist():
            shards[r].search_device(qd.data_ptr() + qi * d * 4, 1, k, 0, 0, mine.data_ptr() + qi * kp * 8, id_base=bounds[r], stream=st)
        lists[:, r, :] = mine
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    N.merge_keys_device(0, lists.data_ptr(), nq, world, k, 0, sc.data_ptr(), ids.data_ptr(), st)
    torch.cuda.synchronize()
    out["batch_equal"] = bool(np.array_equal(ids.cpu().numpy(), Iw) and np.array_equal(sc.cpu().numpy(), Dw))
    out["batch_took_tensor_path"] = bool(took)

    def by_id(ixs, bnds, gid, kk):
        owner = int(np.searchsorted(np.asarray(bnds[1:]), gid, side="right"))
        row = torch.from_numpy(ixs[owner].reconstruct(gid - bnds[owner])).to(dev)
        kq = N.kpad(kk + 1)
        lst = torch.zeros((1, len(ixs), kq), dtype=torch.int64, device=dev)
        for r, ix in enumerate(ixs):
            mine = torch.zeros((1, kq), dtype=torch.int64, device=dev)
            ix.search_device(row.data_ptr(), 1, kk + 1, 0, 0, mine.data_ptr(), id_base=bnds[r], stream=st)
            lst[0, r] = mine[0]
        s1 = torch.empty((1, kk + 1), dtype=torch.float32, device=dev)
        i1 = torch.empty((1, kk + 1), dtype=torch.int64, device=dev)
        N.merge_keys_device(0, lst.data_ptr(), 1, len(ixs), kk + 1, 0, s1.data_ptr(), i1.data_ptr(), st)
        torch.cuda.synchronize()
        keep = (i1[0] != gid).cpu().numpy()
        return s1.cpu().numpy()[0][keep][:kk], i1.cpu().numpy()[0][keep][:kk]

    ok = True
    for gid in (5, bounds[1], bounds[2] - 1, n - 7, int(rng.integers(0, n))):  # (c)
        s, ids = by_id(shards, bounds, int(gid), k)
        De, Ie = whole.search(x[gid], k + 1)
        keep = Ie[0] != gid
        ok &= bool(np.array_equal(ids, Ie[0][keep][:k]) and np.array_equal(s, De[0][keep][:k]))
    out["by_id_equal"] = ok
    whole.close()
    for ix in shards:
        ix.close()
    # (d) config-5 miniature on bf16 shards
    n5, d5, nq5 = 300_000, 768, 64
    x5, rng5 = corpus(n5, d5, 9)
    b5 = shard_bounds(n5, world)
    sh5 = []
    for r in range(world):
        ix = N.NativeIndex(d5, 0, N.STORE_BF16, 0)
        ix.add(x5[b5[r]: b5[r + 1]])
        sh5.append(ix)
    ids5 = rng5.integers(0, n5, nq5)
    got = [by_id(sh5, b5, int(g), k)[1] for g in ids5.tolist()]
    c_oracle.set_threads(c_oracle.host_cores())
    # the stored (bf16-rounded) row is the query, as in the sharded path: fp32 ground truth uses the fp32 row
    _, Io = c_oracle.search(x5, x5[ids5], k + 1, nthreads=c_oracle.host_cores())
    want = [row[row != gid][:k] for row, gid in zip(Io, ids5)]
    out["config5_recall_at_100_vs_fp32_oracle"] = recall(got, want)
    out["config5_queries"] = nq5
    for ix in sh5:
        ix.close()
    print("RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        real_ranks()
    else:
        emulated()
