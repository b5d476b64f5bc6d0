"""Randomised differential test: the CUDA path through the C ABI against the numpy oracle over random
shapes, k, metrics, storage types, predicates, shard splits and launch geometries.  Seeds are fixed, so a
failure reproduces; every case is small enough for the oracle to finish in well under a second.  Rows have
norms of order one (as everything the reference stores): with scores of magnitude d the north-star tolerance
1e-5 |s| + 1e-6 is not meaningful for the near-zero scores of a k = n search."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import flat_ip as O
from tests.conftest import has_gpu
from tests.test_gpu_parity import ATOL, RTOL, check_against_oracle, make_index, make_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_gpu():
        pytest.skip("no GPU")


def N():
    from photo_search_engine_b200 import _native

    return _native


DIMS = [1, 3, 8, 20, 64, 100, 128, 200, 256, 384, 512, 700, 768, 1024, 1030, 1536, 2048, 3000, 4096]


def _rows(rng, n, d, kind):
    x = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "raw":  # not normalised: row norms spread over 0.5 .. 2
        x *= (rng.uniform(0.5, 2.0, (n, 1)) / np.sqrt(d)).astype(np.float32)
    elif kind == "unit":
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-20)
    elif kind == "clustered":
        c = rng.standard_normal((7, d)).astype(np.float32)
        x = c[rng.integers(0, 7, n)] + 0.05 * x
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-20)
    elif kind == "dupes":  # many exact duplicates: ties must order by id
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-20)
        x = x[rng.integers(0, max(2, n // 8), n)]
    return np.ascontiguousarray(x, np.float32)


@pytest.mark.parametrize("seed", range(36))
def test_random_case_against_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice(DIMS))
    n = int(rng.integers(1, 40_000 if d <= 1024 else 6_000))
    metric = int(rng.random() < 0.25)        # 1 = L2
    dtype = int(rng.random() < 0.25)         # 1 = bf16 storage
    kind = str(rng.choice(["unit", "raw", "clustered", "dupes"]))
    x = _rows(rng, n, d, kind)
    nq = int(rng.integers(1, 4))
    q = _rows(rng, nq, d, "unit" if kind != "raw" else "raw")
    if n > 3:
        q[0] = x[int(rng.integers(0, n))]
    ix = make_index(x, metric=metric, dtype=dtype, chunk=int(rng.integers(1, n + 1)) if rng.random() < 0.3 else None)
    stored = x
    if dtype == 1:  # the oracle sees the rows as the device holds them
        import torch

        stored = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    oracle = make_oracle(stored, metric=metric)
    for key, val in (("warps", int(rng.choice([2, 4, 8, 16]))), ("stages", int(rng.choice([0, 2, 3]))),
                     ("deal", int(rng.random() < 0.8)), ("dyn_tail", int(rng.random() < 0.8)),
                     ("static_batch", int(rng.choice([0, 1, 3, 8, 32]))), ("filter_mode", int(rng.choice([0, 1, 2, 3])))):
        ix.set_tunable(key, val)
    mask = flt = None
    if rng.random() < 0.6:
        words = rng.integers(0, 1000, n).astype(np.uint64)  # dt = 0: no datetime, fails a range clause
        ix.set_attrs(0, words)
        lo, hi = sorted(int(v) for v in rng.integers(0, 1100, 2))
        lo = max(lo, 1)
        hi = max(hi, lo)
        flt = N().PsxFilter(flags=N().F_NEED_DT | N().F_START | N().F_END, start=lo, end=hi)
        mask = (words >= lo) & (words <= hi)
    for k in sorted({1, int(rng.integers(1, 64)), int(rng.integers(1, 3000)), n}):
        D, I = ix.search(q, k, flt)
        assert D.shape == (nq, k)
        check_against_oracle(D, I, oracle, q, k, mask=mask)
    ix.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_shard_split_equals_whole(seed):
    """Rows split at random points into 2-4 shards, each scanned with its id_base (with and without a
    predicate), the key lists merged on the device: bit-identical to the unsharded scan."""
    import torch

    rng = np.random.default_rng(50 + seed)
    d = int(rng.choice([64, 200, 768, 1024]))
    n = int(rng.integers(5_000, 30_000))
    k = int(rng.choice([1, 10, 100, 500]))
    x = _rows(rng, n, d, str(rng.choice(["unit", "dupes"])))
    q = _rows(rng, 2, d, "unit")
    words = rng.integers(1, 1000, n).astype(np.uint64)
    flt = N().PsxFilter(flags=N().F_NEED_DT | N().F_END, end=int(rng.integers(100, 900))) if rng.random() < 0.6 else None
    whole = make_index(x)
    whole.set_attrs(0, words)
    Dw, Iw = whole.search(q, k, flt)
    cuts = sorted({0, n, *[int(v) for v in rng.integers(1, n, int(rng.integers(1, 4)))]})
    kp = N().kpad(k)
    keys = torch.zeros((2, len(cuts) - 1, kp), dtype=torch.int64, device="cuda")
    qd = torch.from_numpy(q).cuda()
    st = torch.cuda.current_stream().cuda_stream
    shards = []
    for si, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        sh = make_index(x[a:b])
        sh.set_attrs(0, words[a:b])
        shards.append(sh)
        mine = torch.zeros((2, kp), dtype=torch.int64, device="cuda")
        sh.search_device(qd.data_ptr(), 2, k, 0, 0, mine.data_ptr(), flt=flt, id_base=a, stream=st)
        keys[:, si, :] = mine
    sc = torch.empty((2, k), device="cuda")
    ids = torch.empty((2, k), dtype=torch.int64, device="cuda")
    N().merge_keys_device(0, keys.data_ptr(), 2, len(cuts) - 1, k, 0, sc.data_ptr(), ids.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy(), Iw) and np.array_equal(sc.cpu().numpy(), Dw)
    for sh in shards:
        sh.close()
    whole.close()
