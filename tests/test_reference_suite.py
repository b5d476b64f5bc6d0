"""Run the reference's own, unmodified hot-path test files against the drop-in class.

Only possible where the reference checkout exists (the build container); on the GPU box the
same cases run as restated tests in test_gpu_vector_store.py.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import pytest

from tests.conftest import REFERENCE, ROOT, has_gpu

FILES = ["tests/test_vector_store.py", "tests/test_searcher.py"]


def _run(backend: str):
    # the plugin is imported as a top-level module so that ``tests`` resolves to the reference's package
    path = os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")])
    env = dict(os.environ, PSX_REF_BACKEND=backend, PYTHONPATH=path, PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "ref_inject_plugin", "-p", "no:cacheprovider",
           "--rootdir", REFERENCE, *[os.path.join(REFERENCE, f) for f in FILES]]
    return subprocess.run(cmd, cwd=tempfile.gettempdir(), env=env, capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")), reason="reference checkout not present")
def test_reference_tests_pass_on_host_logic():
    proc = _run("fake")
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    assert "52 passed" in proc.stdout, tail


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "tests")), reason="reference checkout not present")
def test_reference_tests_pass_on_gpu():
    if not has_gpu():
        pytest.skip("no GPU")
    proc = _run("gpu")
    assert proc.returncode == 0, (proc.stdout + proc.stderr)[-3000:]
