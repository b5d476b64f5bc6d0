"""Row shards across ranks under the driver: ``ShardedIndex.search``, the sharded query
batch, ``search_by_id`` and a config-5 miniature (bf16 shards, recall@100 against the fp32 CPU oracle) run as a
``torchrun`` job with one rank per GPU when the box has at least two GPUs, and with the ranks emulated on device 0
through the same C-ABI calls when it has one.  tests/_torchrun_case.py is the body."""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys

import pytest

from tests.conftest import ROOT, has_gpu

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_paths_equal_one_index():
    if not has_gpu():
        pytest.skip("no GPU")
    import torch

    ngpu = torch.cuda.device_count()
    case = os.path.join(ROOT, "tests", "_torchrun_case.py")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if ngpu >= 2:
        world = min(ngpu, 4)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), case]
    else:
        world = 3
        env.pop("WORLD_SIZE", None)
        cmd = [sys.executable, case]
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, tail
    out = json.loads(lines[-1][len("RESULT "):])
    assert out["world"] == world
    assert out["single_equal"], out
    assert out["tie_order"][0] == 11 and out["tie_order"][1] == 290_000 - 7   # exact tie across shards: lower id first
    assert out["batch_equal"] and out["batch_took_tensor_path"], out
    assert out["by_id_equal"], out
    # bf16 storage is approximate: recall, not identity (uniform random unit vectors are its worst case)
    assert out["config5_queries"] >= 64 and out["config5_recall_at_100_vs_fp32_oracle"] >= 0.97, out
