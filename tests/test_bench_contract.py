"""bench.py keeps the driver's contract: ONE JSON line with the agreed keys, for both arms.  The reference arm runs on
the CPU (no GPU needed); the repository arm is exercised at a reduced row count on a GPU box."""
from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT, has_gpu

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
          "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1")  # as torchrun sets it
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, capture_output=True, text=True, timeout=timeout)
    assert proc.returncode == 0, (proc.stdout + proc.stderr)[-3000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout[-2000:]           # exactly one line on stdout
    return json.loads(lines[0])


def test_reference_arm_line_and_thread_pinning():
    line = _run(["--impl", "reference", "--rows", "60000", "--dim", "256", "--steps", "3", "--warmup", "1"])
    assert COMMON <= set(line) and line["impl"] == "reference"
    assert line["unit"] == "queries/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["value"] == line["value"]
    # the arm uses every host core although OMP_NUM_THREADS=1 is exported (torchrun does that for N > 1)
    assert base["cores"] == max(1, len(os.sched_getaffinity(0)))
    assert base["single_thread_value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = line["config"]
    assert cfg["rows_per_step"] == 60000 and cfg["row_scale"] == 1.0       # measured on the full (small) corpus, nothing extrapolated
    assert abs(line["ms_per_step"] - cfg["ms_per_query_full_corpus"]) < 1e-9
    assert "60000x256 fp32 flat-IP top-100" in cfg["workload"]


@pytest.mark.gpu
def test_repository_arm_line_on_a_small_corpus():
    if not has_gpu():
        pytest.skip("no GPU")
    line = _run(["--rows", "400000", "--dim", "256", "--steps", "8", "--warmup", "3", "--no-extras"])
    assert COMMON | {"roofline", "gpu_launches", "clocks", "parity_spot_check"} <= set(line)
    assert line["n_gpus"] == 1 and line["gpu_launches"] == 8          # one kernel per query: merge and selection are fused in
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["achieved"] > 0 and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["algorithmic_bytes_per_launch"] == 400000 * 256 * 4
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 256 * 4 and e2e["d2h_bytes_per_step"] == 100 * 12
    spot = line["parity_spot_check"]
    assert spot["ids_equal_frac"] == 1.0 and spot["max_rel_score_err"] < 1e-5
    assert line["cpu_baseline"]["cores"] == max(1, len(os.sched_getaffinity(0)))
    assert line["config"]["workload"].startswith("400000x256 fp32 flat-IP top-100")
    assert "value_with_default_pdl1" in line["config"]
