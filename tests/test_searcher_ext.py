"""the opt-in ``BatchedExpansionMixin`` in front of the reference's unmodified
``core.searcher.Searcher`` answers the expansion loop's searches from ONE ``search_batch`` and returns
exactly what the unbatched Searcher returns.  Needs the reference checkout (build container)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import pytest

from tests.conftest import REFERENCE, ROOT, has_gpu


def _run(backend: str):
    path = os.pathsep.join([REFERENCE, os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")])
    env = dict(os.environ, PSX_REF_BACKEND=backend, PYTHONPATH=path, PYTHONDONTWRITEBYTECODE="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_searcher_ext_case.py")], cwd=tempfile.gettempdir(), env=env,
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, (proc.stdout + proc.stderr)[-3000:]
    line = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def _check(out):
    plain, batched = out["plain"], out["batched"]
    # identical results, through the direct call and through Searcher.search(..., "high_recall")
    assert batched["direct"] == plain["direct"] and len(plain["direct"]) > 0
    assert batched["full"] == plain["full"] and len(plain["full"]) > 0
    assert plain["alternatives_run"] == batched["alternatives_run"] == 4
    # the unbatched loop: one search + one embedding call per alternative
    assert plain["after_direct"] == [4, 0, 4, 0]
    # batched: the 3 distinct alternative texts are embedded by ONE batch call and searched by ONE search_batch;
    # the loop's own 4 searches never reach the store
    assert batched["after_direct"] == [0, 1, 0, 1]
    assert batched["stats"]["batches"] >= 1 and batched["stats"]["batched_queries"] >= 3
    assert batched["stats"]["served_from_batch"] >= 4
    assert plain["expansion_triggered_full"] == batched["expansion_triggered_full"]
    # FusedPrefilterMixin: every returned photo passes the filter in both variants; the pre-filtered recall contains
    # every passing photo the reference found, finds the true 10 best passing rows, and is never smaller
    pf = out["prefilter_case"]
    truth = pf["truth"]
    got_plain = [p for p, _ in pf["plain"]]
    got_pre = [p for p, _ in pf["prefilter"]]
    assert set(got_plain) <= set(truth) | set(got_pre)
    assert set(got_pre) == set(truth) and len(got_pre) == 10
    assert len(got_plain) <= len(got_pre)
    # FusedRecallMixin: identical rounds -- results, ranks, buckets, summaries and round quality --
    # with the candidates kept as arrays; only the round with media terms takes the reference path
    rc = out["recall_case"]
    assert len(rc["plain"]) == len(rc["recall"]) == 9
    for i, (a, b) in enumerate(zip(rc["plain"], rc["recall"])):
        assert a["results"] == b["results"], i
        assert a["quality"] == b["quality"], i
    assert any(len(r["results"]) > 0 for r in rc["plain"][2:7])            # the filtered rounds do return photos
    assert rc["recall_stats"]["array_rounds"] == 24 and rc["recall_stats"]["reference_rounds"] == 1
    assert rc["ms_per_round"]["recall"] < rc["ms_per_round"]["plain"]        # and the Python tail got shorter
    assert rc["vector_scores"] is True                                       # scores computed as one array, bit-identical (probe)
    assert rc["odd"]["recall_vector_scores"] is False and rc["odd"]["plain"] == rc["odd"]["recall"] and len(rc["odd"]["plain"]) > 0
    # the same on the Elasticsearch branch: vector hits + keyword hits fused (core/searcher.py:855-988), stale ES documents
    # dropped, duplicate paths collapsed, ES filters applied -- identical results and round quality
    hy = out["hybrid_case"]
    assert len(hy["plain"]) == len(hy["recall"]) == 4
    for i, (a, b) in enumerate(zip(hy["plain"], hy["recall"])):
        assert a["results"] == b["results"] and len(a["results"]) > 0, i
        assert a["quality"] == b["quality"], i
    assert any(r[3] > 0 for rnd in hy["plain"] for r in rnd["results"])        # keyword scores really took part
    assert hy["recall_stats"]["array_rounds"] == 4


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "core")), reason="reference checkout not present")
def test_batched_expansion_equals_reference_loop_on_host_logic():
    _check(_run("fake"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "core")), reason="reference checkout not present")
def test_batched_expansion_equals_reference_loop_on_gpu():
    if not has_gpu():
        pytest.skip("no GPU")
    _check(_run("gpu"))
