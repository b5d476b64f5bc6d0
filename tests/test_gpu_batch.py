"""K3 (tensor-core batched search) parity: results through the tcgen05 TF32 GEMM + fused selection +
exact re-score must be BIT-IDENTICAL to the single-query streaming scan (same per-row reduction
tree), which itself is checked against the oracle in test_gpu_parity.py.  Queries whose proof
obligation fails are re-run on the scan by psx_search, so equality must hold for every query."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import flat_ip as O
from tests.conftest import has_gpu
from tests.test_gpu_parity import check_against_oracle, make_oracle, unit_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not has_gpu():
        pytest.skip("no GPU")


def N():
    from photo_search_engine_b200 import _native

    return _native


def _both_paths(ix, q, k):
    ix.set_tunable("batch_min", 0)  # streaming scan, one launch per query
    Ds, Is = ix.search(q, k)
    ix.set_tunable("batch_min", 2)  # tensor-core path
    before = ix.batch_stats()
    Db, Ib = ix.search(q, k)
    after = ix.batch_stats()
    return (Ds, Is), (Db, Ib), (after[0] - before[0], after[1] - before[1])


@pytest.mark.parametrize("n,d,nq,k", [(70_000, 64, 5, 10), (200_000, 256, 128, 100), (150_000, 1024, 200, 100),
                                      (100_000, 768, 300, 50), (66_000, 100, 17, 100), (131_072, 4096, 9, 20), (110_000, 64, 40, 512)])
def test_batch_equals_scan(n, d, nq, k):
    rng = np.random.default_rng(n + d + nq)
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    q[0] = x[12345]                       # an exact self-match
    q[1] = (x[777] + 0.1 * q[1]).astype(np.float32)
    q[1] /= np.linalg.norm(q[1])
    ix = N().NativeIndex(d)
    ix.add(x)
    (Ds, Is), (Db, Ib), (served, fallbacks) = _both_paths(ix, q, k)
    assert served == nq, "the batch did not go through the tensor-core path"
    assert np.array_equal(Ib, Is)
    assert np.array_equal(Db, Ds)         # bit-identical scores
    assert Ib[0, 0] == 12345 and Ib[1, 0] == 777
    assert fallbacks <= nq // 10, f"{fallbacks} of {nq} queries needed the fallback"
    # and the scan agrees with the oracle on a few of them
    oracle = make_oracle(x)
    check_against_oracle(Ds[:3], Is[:3], oracle, q[:3], k)
    ix.close()


@pytest.mark.parametrize("k", [675, 850, 1025, 1333, 2048])
def test_batch_serves_the_call_site_k(k):
    """candidate_k of the reference's expansion / reflection rounds (core/searcher.py:771-820: 675, 850, 1025, up to
    1333) and the pass maximum go through the tensor-core path -- not the nq-scans fallback -- and stay bit-identical."""
    rng = np.random.default_rng(k)
    n, d, nq = 240_000, 128, 12          # the path asks for n >= 24 * (4k + 64) rows
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    q[0] = x[31337]
    ix = N().NativeIndex(d)
    ix.add(x)
    (Ds, Is), (Db, Ib), (served, fallbacks) = _both_paths(ix, q, k)
    assert served == nq, "k above 512 fell back to sequential scans"
    assert np.array_equal(Ib, Is) and np.array_equal(Db, Ds)
    assert Ib[0, 0] == 31337
    assert fallbacks <= 2, fallbacks
    ix.close()


def test_batch_with_non_unit_queries():
    """The rounding bound of the certificate scales with |q|: it is computed per query on the device, so callers
    (ShardedIndex, psx_search_batch_device users) need not supply norms.  Queries of norm 0.01 .. 30 stay exact."""
    rng = np.random.default_rng(21)
    n, d, nq, k = 100_000, 256, 16, 50
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d) * np.logspace(-2, 1.5, nq).astype(np.float32)[:, None]
    ix = N().NativeIndex(d)
    ix.add(x)
    (Ds, Is), (Db, Ib), (served, fallbacks) = _both_paths(ix, q, k)
    assert served == nq and np.array_equal(Ib, Is) and np.array_equal(Db, Ds)
    assert fallbacks <= 2
    ix.close()


def test_bf16_certificate_survives_sign_aligned_rounding():
    """Adversarial for the bf16 GEMM certificate: q and the rows A all sit 0.498 ulp above a bf16 grid
    point, so BOTH operands round down and the bf16 score of A is 2*2^-8 = 0.78 % below its exact score (1.0000
    vs 1.0078).  Decoy rows B (exact == bf16 score 1.0068) and a filler band F (1.0009 .. 1.0019, 3000 rows, where
    the threshold lands) rank above A in bf16 arithmetic although A is the true top-100.  A rounding bound of
    2^-8 |q||x| (one operand) certifies the wrong answer B; the rigorous 2*2^-8 bound must refuse and fall back."""
    rng = np.random.default_rng(5)
    n, d, nq, k = 70_000, 1024, 8, 100
    g = np.float32(2.0 ** -5)
    down, up = np.float32(1 + 0.498 * 2.0 ** -7), np.float32(1 + 0.502 * 2.0 ** -7)
    x = unit_rows(rng, n, d)
    special = rng.permutation(n)[: 100 + 120 + 3000]
    rows_a, rows_b, rows_f = special[:100], special[100:220], special[220:]

    def row(value, nnz):
        v = np.full(d, value, np.float32)
        v[rng.permutation(d)[: d - nnz]] = 0.0
        return v

    for r in rows_a:
        x[r] = row(g * down, 1024)
    for r in rows_b:
        x[r] = row(g * up, 1023)
    for r in rows_f:
        x[r] = row(g * up, 1017 + int(rng.integers(0, 2)))
    q = np.tile(np.full(d, g * down, np.float32), (nq, 1))
    exact = N().NativeIndex(d)
    exact.add(x)
    exact.set_tunable("batch_min", 0)
    De, Ie = exact.search(q, k)
    exact.close()
    assert set(Ie[0].tolist()) == set(rows_a.tolist())       # the true top-100 is A
    ix = N().NativeIndex(d, 0, N().STORE_BF16_MASTER, 0)
    ix.add(x)
    ix.set_tunable("batch_min", 2)
    ix.set_tunable("batch_bf16", 1)
    before = ix.batch_stats()
    Db, Ib = ix.search(q, k)
    served, fallbacks = (a - b for a, b in zip(ix.batch_stats(), before))
    assert served == nq
    assert np.array_equal(Ib, Ie) and np.array_equal(Db, De)
    assert fallbacks == nq                                    # no query may be certified from the bf16 scores here
    ix.close()


@pytest.mark.parametrize("n,d,nq,k", [(150_000, 1024, 200, 100), (100_000, 768, 300, 50), (70_000, 96, 129, 10)])
def test_cta_pair_kernel_equals_scan(n, d, nq, k):
    """The cta_group::2 variant (two SMs share one 256 x 256 tile, M = 256 MMAs issued by the leader CTA)."""
    rng = np.random.default_rng(n + d + nq + 1)
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    q[0] = x[4321]
    q[nq - 1] = x[99]                     # a query handled by the second CTA of the pair
    ix = N().NativeIndex(d)
    ix.add(x)
    ix.set_tunable("batch_pair", 1)
    (Ds, Is), (Db, Ib), (served, fallbacks) = _both_paths(ix, q, k)
    assert served == nq
    assert np.array_equal(Ib, Is) and np.array_equal(Db, Ds)
    assert Ib[0, 0] == 4321 and Ib[nq - 1, 0] == 99
    assert fallbacks <= nq // 10
    ix.close()


@pytest.mark.parametrize("nq", [40, 200])
def test_batch_with_fused_predicate(nq):
    """The EXIF predicate in the batched path (candidate epilogue + threshold sample) == the filtered scan."""
    rng = np.random.default_rng(nq)
    n, d, k = 120_000, 256, 100
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    ix = N().NativeIndex(d)
    ix.add(x)
    ix.set_attrs(0, np.arange(n, dtype=np.uint64) + np.uint64(1))   # dt = 1 + row
    for lo, hi in [(30_000, 90_000), (1, 3_000), (100_000, 100_050)]:  # 50 %, 2.5 %, fewer rows than k
        flt = N().PsxFilter(flags=N().F_NEED_DT | N().F_START | N().F_END, start=lo, end=hi)
        ix.set_tunable("batch_min", 0)
        Ds, Is = ix.search(q, k, flt)
        ix.set_tunable("batch_min", 2)
        before = ix.batch_stats()
        Db, Ib = ix.search(q, k, flt)
        assert ix.batch_stats()[0] - before[0] == nq
        assert np.array_equal(Ib, Is) and np.array_equal(Db, Ds), (lo, hi)
        valid = Ib[Ib >= 0]
        assert ((valid >= lo - 1) & (valid < hi)).all()
    ix.close()


def test_batch_on_clustered_data_with_duplicates():
    """Tight clusters + exact duplicates: tiny score gaps make the TF32 proof fail for some
    queries; the fallback must keep every answer exact and ties ordered by id."""
    rng = np.random.default_rng(3)
    n, d, nq, k = 120_000, 128, 64, 100
    centers = unit_rows(rng, 50, d)
    x = centers[rng.integers(0, 50, n)] + 0.01 * rng.standard_normal((n, d)).astype(np.float32)
    x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    x[5000:5200] = x[100]                 # 200 exact copies
    q = centers[:nq % 50 + 14].repeat(5, axis=0)[:nq].astype(np.float32)
    q[0] = x[100]
    ix = N().NativeIndex(d)
    ix.add(x)
    (Ds, Is), (Db, Ib), (served, fallbacks) = _both_paths(ix, q, k)
    assert np.array_equal(Ib, Is) and np.array_equal(Db, Ds)
    assert Ib[0, 0] == 100 and Ib[0, 1] == 5000 and (np.diff(Ib[0, 1:100]) > 0).all()
    ix.close()


def test_batch_device_api_flags():
    import torch

    rng = np.random.default_rng(9)
    n, d, nq, k = 100_000, 512, 40, 100
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    ix = N().NativeIndex(d)
    ix.add(x)
    ix.set_tunable("batch_min", 0)
    Ds, Is = ix.search(q, k)
    qd = torch.from_numpy(q).cuda()
    sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    flags = torch.full((nq,), -1, dtype=torch.int32, device="cuda")
    keys = torch.zeros((nq, N().kpad(k)), dtype=torch.int64, device="cuda")
    ix.search_batch_device(qd.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), flags.data_ptr(), keys.data_ptr(),
                           qnorm_max=1.0, id_base=1000, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    f = flags.cpu().numpy()
    assert (f >= 0).all()
    ok = f == 0
    assert ok.mean() > 0.9
    assert np.array_equal(ids.cpu().numpy()[ok], Is[ok] + 1000)
    assert np.array_equal(sc.cpu().numpy()[ok], Ds[ok])
    ix.close()


@pytest.mark.parametrize("n,d,nq,k", [(150_000, 1024, 200, 100), (100_000, 768, 64, 50), (70_000, 100, 256, 10), (90_000, 200, 5, 100)])
def test_bf16_master_batch_uses_bf16_gemm_and_stays_exact(n, d, nq, k):
    """bf16 + fp32 master index: the batched GEMM streams the bf16 rows (kind::f16, operands rounded to 8 bits),
    the survivors are re-scored on the fp32 master -- every certified result is bit-identical to an fp32 index,
    with the bf16 GEMM (default) and with the TF32 GEMM over the master (batch_bf16 = 0)."""
    rng = np.random.default_rng(n + d + nq + 7)
    x = unit_rows(rng, n, d)
    q = unit_rows(rng, nq, d)
    q[0] = x[4242]
    exact = N().NativeIndex(d)
    exact.add(x)
    exact.set_tunable("batch_min", 0)
    De, Ie = exact.search(q, k)
    exact.close()
    ix = N().NativeIndex(d, 0, N().STORE_BF16_MASTER, 0)
    ix.add(x)
    for bf in (1, 0):
        ix.set_tunable("batch_bf16", bf)
        ix.set_tunable("batch_min", 2)
        before = ix.batch_stats()
        Db, Ib = ix.search(q, k)
        served, fallbacks = (a - b for a, b in zip(ix.batch_stats(), before))
        assert served == nq
        assert np.array_equal(Ib, Ie) and np.array_equal(Db, De), bf
        assert Ib[0, 0] == 4242
        assert fallbacks <= max(1, nq // 10), (bf, fallbacks)
    ix.close()
