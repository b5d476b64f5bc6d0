"""One handle, several devices, ONE process (``psx_create_sharded`` / ``VectorStore(devices=[...])``).

The reference constructs a single ``VectorStore`` in a single process (main.py:59-68), so the 8-GPU row sharding has
to live behind that one instance.  Every result must be bit-identical to a single-device index -- ids, scores and
tie order -- whatever path a query takes: the fused NVLink exchange (single queries), event-ordered key lists
(query batches on the tensor cores, k > 2048 paging, the bf16 + fp32-master tier), the fused exchange's time-out path.

On a one-GPU box the shards are several entries of device 0 (``devices=[0, 0, 0]``): the same code path -- one child
index, stream and scratch per entry, peer-addressed receive buffer, flags, merge kernel -- without NVLink in between.
With more GPUs the real ordinals are used as well.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from tests.conftest import REFERENCE, ROOT, has_gpu
from tests.test_gpu_parity import check_against_oracle, make_oracle, unit_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_gpu():
        pytest.skip("no GPU")


def N():
    from photo_search_engine_b200 import _native

    return _native


def device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 4))))
    if n >= 8:
        lists.append(list(range(8)))
    return lists


def pair(d, devices, metric=0, dtype=0, min_rows=1):
    one = N().NativeIndex(d, metric, dtype, 0)
    many = N().NativeIndex(d, metric, dtype, devices)
    many.set_tunable("shard_min_rows", min_rows)
    return one, many


@pytest.mark.parametrize("devices", device_lists())
@pytest.mark.parametrize("n,d,k", [(5000, 64, 10), (40_000, 1024, 100), (20_001, 100, 1333), (3000, 4096, 50), (9, 8, 5)])
def test_single_queries_equal_one_device(devices, n, d, k):
    rng = np.random.default_rng(n + d + len(devices))
    x = unit_rows(rng, n, d)
    x[n // 2] = x[1]                      # an exact tie living on different shards
    q = unit_rows(rng, 3, d)
    q[0] = x[1]
    one, many = pair(d, devices)
    one.add(x)
    many.add(x)
    for qi in range(3):                   # nq = 1: the fused exchange
        Do, Io = one.search(q[qi], k)
        Dm, Im = many.search(q[qi], k)
        assert np.array_equal(Im, Io) and np.array_equal(Dm, Do), (devices, qi)
        if qi == 0:  # the query is row 1, whose exact copy sits at n // 2 on another shard: lower id first
            assert Im[0, 0] == 1 and (min(k, n) < 2 or Im[0, 1] == n // 2)
    rows = many.shard_rows()
    assert sum(r for _, r in rows) == n and [dv for dv, _ in rows] == devices
    if n >= 4 * len(devices):
        assert min(r for _, r in rows) > 0, rows   # the corpus really is spread over every shard
        assert many.group_stats()[0] >= 3          # and the queries went through the fused exchange
    kk = min(k, n)
    check_against_oracle(Dm[:, :kk], Im[:, :kk], make_oracle(x), q[2:3], kk)
    one.close()
    many.close()


@pytest.mark.parametrize("devices", device_lists()[:2])
def test_l2_bf16_and_mixed_tiers(devices):
    rng = np.random.default_rng(17)
    n, d, k = 30_000, 256, 64
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, 2, d)
    for metric, dtype in ((1, N().STORE_F32), (0, N().STORE_BF16), (0, N().STORE_BF16_MASTER), (1, N().STORE_BF16)):
        one, many = pair(d, devices, metric, dtype)
        one.add(x)
        many.add(x)
        for qi in range(2):
            Do, Io = one.search(q[qi], k)
            Dm, Im = many.search(q[qi], k)
            assert np.array_equal(Im, Io) and np.array_equal(Dm, Do), (metric, dtype)
        one.close()
        many.close()


@pytest.mark.parametrize("devices", device_lists())
def test_predicate_paging_and_small_batches(devices):
    rng = np.random.default_rng(23)
    n, d = 50_000, 128
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, 6, d)
    one, many = pair(d, devices)
    one.add(x)
    many.add(x)
    words = np.arange(n, dtype=np.uint64) + np.uint64(1)
    one.set_attrs(0, words)
    many.set_attrs(0, words)
    # EXIF predicate (a window that spans two shards, and one with fewer rows than k)
    for lo, hi in ((n // len(devices) - 500, n // len(devices) + 800), (10, 40)):
        flt = N().PsxFilter(flags=N().F_NEED_DT | N().F_START | N().F_END, start=lo, end=hi)
        Do, Io = one.search(q[0], 100, flt)
        Dm, Im = many.search(q[0], 100, flt)
        assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
        valid = Im[Im >= 0]
        assert ((valid >= lo - 1) & (valid < hi)).all()
    # k above one pass: paged on every shard, page ceilings read from the merging device
    Do, Io = one.search(q[1], 5000)
    Dm, Im = many.search(q[1], 5000)
    assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
    # k beyond the corpus: padded
    Dm, Im = many.search(q[1][None], n + 7)
    assert (Im[0, n:] == -1).all() and len(set(Im[0, :n].tolist())) == n
    # several queries in one call: below / above the batch threshold (small corpus: per-shard scans into key lists)
    for nq in (2, 6):
        Do, Io = one.search(q[:nq], 50)
        Dm, Im = many.search(q[:nq], 50)
        assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
    one.close()
    many.close()


@pytest.mark.parametrize("devices", device_lists()[:3])
def test_batches_take_the_tensor_core_path_per_shard(devices):
    rng = np.random.default_rng(31)
    G = len(devices)
    n, d, nq, k = 70_000 * G, 128, 40, 100
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    q[0] = x[n - 5]
    one, many = pair(d, devices)
    one.add(x)
    many.add(x)
    one.set_tunable("batch_min", 0)
    Do, Io = one.search(q, k)
    before = many.batch_stats()
    Dm, Im = many.search(q, k)
    served, fallbacks = (a - b for a, b in zip(many.batch_stats(), before))
    assert served == nq, "the shards did not take the tensor-core path"
    assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
    assert Im[0, 0] == n - 5
    assert fallbacks <= nq * G // 10
    one.close()
    many.close()


@pytest.mark.parametrize("dtype", ["fp32", "bf16+fp32"])
def test_appends_rebalance_and_keep_ids(dtype):
    """add_item-style growth (core/indexer.py:858): rows trickle in, the tail shard grows, the layout is re-split with
    device-to-device copies -- ids, reconstruct, read_rows and attribute words follow their rows (and, on the
    bf16 + fp32-master tier, both copies of every row)."""
    rng = np.random.default_rng(41)
    d, total = 32, 6000
    x = unit_rows(rng, total, d)
    one, many = pair(d, [0, 0, 0], dtype=N().STORE_F32 if dtype == "fp32" else N().STORE_BF16_MASTER, min_rows=64)
    words = (np.arange(total, dtype=np.uint64) % np.uint64(7)) + np.uint64(1)
    done = 0
    layouts = set()
    for step, m in enumerate([1, 1, 5, 60, 200, 33, 700, 1, 1500, 999, 2500]):
        m = min(m, total - done)
        one.add(x[done: done + m])
        many.add(x[done: done + m])
        done += m
        many.sync()
        one.sync()
        many.set_attrs(0, words[:done])
        one.set_attrs(0, words[:done])
        rows = tuple(r for _, r in many.shard_rows())
        layouts.add(rows)
        assert sum(rows) == done == many.ntotal
        assert max(rows) <= max(64, 1.25 * (done / 3 + 1)) + 1, rows     # never far from an even split
        q = unit_rows(rng, 1, d)[0]
        k = min(20, done)
        for flt in (None, N().PsxFilter(flags=N().F_NEED_DT | N().F_START | N().F_END, start=3, end=5)):
            Do, Io = one.search(q, k, flt)
            Dm, Im = many.search(q, k, flt)
            assert np.array_equal(Im, Io) and np.array_equal(Dm, Do), (step, rows)
        probe = rng.integers(0, done, 5)
        for i in probe:
            assert np.array_equal(many.reconstruct(int(i)), x[i])
    assert len(layouts) >= 5
    assert np.array_equal(many.read_rows(0, total), x) and np.array_equal(many.read_rows(1999, 2003), x[1999:4002])
    many.reset()
    assert many.ntotal == 0 and all(r == 0 for _, r in many.shard_rows())
    many.add(x[:10])
    assert np.array_equal(many.search(x[3], 1)[1], [[3]])
    one.close()
    many.close()


def test_reserve_then_bulk_load_splits_evenly():
    rng = np.random.default_rng(43)
    n, d = 90_001, 64
    x = unit_rows(rng, n, d)
    many = N().NativeIndex(d, 0, 0, [0, 0, 0, 0])
    many.reserve(n)
    for off in range(0, n, 20_000):
        many.add(x[off: off + 20_000])
    many.sync()
    rows = [r for _, r in many.shard_rows()]
    assert sum(rows) == n and max(rows) - min(rows) <= 4, rows
    one = N().NativeIndex(d)
    one.add(x)
    q = unit_rows(rng, 1, d)[0]
    assert np.array_equal(many.search(q, 100)[1], one.search(q, 100)[1])
    one.close()
    many.close()


def test_bulk_device_ingest_lands_on_the_right_shards():
    import torch

    ndev = torch.cuda.device_count()
    devices = list(range(min(ndev, 4))) if ndev >= 2 else [0, 0]
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    n, d = 64_000, 128
    x = torch.randn((n, d), generator=gen, device="cuda:0")
    x = (x / x.norm(dim=1, keepdim=True)).contiguous()
    one = N().NativeIndex(d)
    many = N().NativeIndex(d, 0, 0, devices)
    many.reserve(n)
    st = torch.cuda.current_stream().cuda_stream
    for off in range(0, n, 24_000):
        m = min(24_000, n - off)
        one.add_device(x[off:].data_ptr(), m, stream=st)
        many.add_device(x[off:].data_ptr(), m, stream=st)
    words = torch.arange(n, device="cuda:0", dtype=torch.int64) % 11 + 1
    one.set_attrs_device(0, words.data_ptr(), n, stream=st)
    many.set_attrs_device(0, words.data_ptr(), n, stream=st)
    rows = [r for _, r in many.shard_rows()]
    assert sum(rows) == n and max(rows) - min(rows) <= len(devices)
    q = x[-3].cpu().numpy()
    flt = N().PsxFilter(flags=N().F_NEED_DT | N().F_END, end=4)
    for f in (None, flt):
        Do, Io = one.search(q, 64, f)
        Dm, Im = many.search(q, 64, f)
        assert np.array_equal(Im, Io) and np.array_equal(Dm, Do)
    assert Im[0, 0] != n - 3 and many.search(q, 1)[1][0, 0] == n - 3
    one.close()
    many.close()


def test_fused_exchange_timeout_is_survivable():
    """A shard that never publishes (fault hook): the merge kernel's bounded wait reports it in the status word, the
    query is answered over the event-ordered path, the context stays healthy and later queries are fused again."""
    rng = np.random.default_rng(47)
    n, d, k = 9000, 64, 25
    x = unit_rows(rng, n, d)
    one, many = pair(d, [0, 0, 0])
    one.add(x)
    many.add(x)
    q = unit_rows(rng, 2, d)
    want = one.search(q[0], k)
    assert np.array_equal(many.search(q[0], k)[1], want[1])
    many.set_tunable("xchg_timeout_ms", 30)
    many.set_tunable("fault_skip_publish", 1)
    got = many.search(q[0], k)
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])
    fused, keyed, timeouts = many.group_stats()
    assert timeouts == 1 and keyed >= 1
    many.set_tunable("fault_skip_publish", -1)
    many.set_tunable("xchg_timeout_ms", 0)
    got = many.search(q[1], k)
    assert np.array_equal(got[1], one.search(q[1], k)[1])
    assert many.group_stats()[2] == 1
    one.close()
    many.close()


def test_device_pointer_calls_are_refused_on_a_group():
    many = N().NativeIndex(16, 0, 0, [0, 0])
    many.add(np.eye(16, dtype=np.float32))
    with pytest.raises(RuntimeError) as err:
        many.storage_device()
    assert "single-device call" in str(err.value)
    many.close()


def test_vector_store_with_devices_round_trips(tmp_path):
    """The drop-in class over two shards: add_item / search / save / load / clear, files byte-identical to one device."""
    from photo_search_engine_b200.vector_store import VectorStore

    rng = np.random.default_rng(53)
    d, n = 24, 300
    rows = rng.standard_normal((n, d)).astype(np.float32)
    stores = {}
    for tag, kw in (("one", {}), ("two", {"devices": [0, 0]})):
        s = VectorStore(d, str(tmp_path / f"{tag}.index"), str(tmp_path / f"{tag}.json"), **kw)
        if tag == "two":
            s.index.set_tunable("shard_min_rows", 16)
        for i, r in enumerate(rows):
            s.add_item(r.tolist(), {"photo_path": f"/p/{i}.jpg", "exif_data": {"datetime": f"2020-0{1 + i % 9}-10T10:00:00"},
                                    "time_info": {"year": 2020, "month": 1 + i % 9, "season": "夏天", "time_period": "上午",
                                                  "datetime_str": f"2020-0{1 + i % 9}-10T10:00:00"}})
        stores[tag] = s
    q = rng.standard_normal(d).astype(np.float32).tolist()
    for kw in ({}, {"constraints": {"start_date": "2020-03-01", "end_date": "2020-05-31"}}):
        a = stores["one"].search(q, 20, **kw)
        b = stores["two"].search(q, 20, **kw)
        assert [h["metadata"]["photo_path"] for h in a] == [h["metadata"]["photo_path"] for h in b] and len(a) > 0
        assert [h["distance"] for h in a] == [h["distance"] for h in b]
    assert min(r for _, r in stores["two"].index.shard_rows()) > 0
    assert stores["two"].get_embedding_by_photo_path("/p/250.jpg") == stores["one"].get_embedding_by_photo_path("/p/250.jpg")
    for s in stores.values():
        s.save()
    assert open(stores["one"].index_path, "rb").read() == open(stores["two"].index_path, "rb").read()
    again = VectorStore(d, stores["two"].index_path, stores["two"].metadata_path, devices=[0, 0, 0])
    assert again.load() and again.get_total_items() == n
    assert [h["metadata"]["photo_path"] for h in again.search(q, 20)] == [h["metadata"]["photo_path"] for h in stores["one"].search(q, 20)]
    again.clear()
    assert again.get_total_items() == 0 and again.search(q, 5) == []


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")), reason="reference suite neither checked out nor staged")
def test_reference_tests_pass_on_two_shards():
    """The reference's own unmodified tests/test_vector_store.py + tests/test_searcher.py with PSX_DEVICES=0,0 (every
    store they create is split over two shards, down to single rows: PSX_SHARD_MIN_ROWS=1)."""
    import torch

    second = 1 if torch.cuda.device_count() >= 2 else 0
    path = os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")])
    env = dict(os.environ, PSX_REF_BACKEND="gpu", PYTHONPATH=path, PYTHONDONTWRITEBYTECODE="1", PSX_DEVICES=f"0,{second}",
               PSX_SHARD_MIN_ROWS="1")
    files = ["tests/test_vector_store.py", "tests/test_searcher.py"]
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "ref_inject_plugin", "-p", "no:cacheprovider", "--rootdir", REFERENCE,
           *[os.path.join(REFERENCE, f) for f in files]]
    proc = subprocess.run(cmd, cwd=tempfile.gettempdir(), env=env, capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    assert "52 passed" in proc.stdout, tail
