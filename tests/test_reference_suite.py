"""Run the reference's own, unmodified hot-path test files against the drop-in class.

In the build container they run from the checkout (/root/reference); on the GPU box from the
unmodified copy that ``__graft_entry__.build()`` staged under oracle/_ref/reference_suite/ (git-ignored,
see oracle/stage_reference.py).  ``test_staged_reference_suite_is_unmodified`` pins the staged files to
the committed SHA-256 digests.
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


def test_staged_reference_suite_is_unmodified():
    """What travels to the GPU box is byte-identical to the reference's files (digests committed under
    tests/golden/; regenerate with ``python -m oracle.stage_reference`` if the reference ever changes)."""
    from oracle import stage_reference as sr

    folder = sr.stage()
    if folder is None:
        pytest.skip("neither a reference checkout nor a staged copy is present")
    pinned = sr.pinned_manifest()
    assert sorted(pinned) == sorted(sr.FILES)
    assert sr.manifest_of(folder) == pinned
    if os.path.isdir(os.path.join(sr.DEFAULT_SOURCE, "tests")):
        assert sr.manifest_of(sr.DEFAULT_SOURCE) == pinned


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
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    assert "52 passed" in proc.stdout, tail
