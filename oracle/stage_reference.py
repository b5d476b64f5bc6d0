"""Stage the reference's own hot-path test files (and the few modules they import) for the GPU box.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container but not on the GPU
box, and the reference is a Python application (no compiled sources), so the "compile the reference
into oracle/_ref/" recipe of a C reference becomes: copy the closed set of files the two suites
import, UNMODIFIED, into ``oracle/_ref/reference_suite/`` -- a directory that is git-ignored (it never
enters history: no reference source is committed) but not gpurun-ignored, so it travels with the
snapshot exactly like the built ``.so`` files.  ``__graft_entry__.build()`` calls :func:`stage`.

The file list is the import closure of

    tests/test_vector_store.py, tests/test_searcher.py      (the suites)
    tests/helpers.py, tests/__init__.py                     (their fakes)
    core/searcher.py, core/__init__.py                      (the caller of the hot path)
    utils/__init__.py, utils/path_utils.py, utils/structured_analysis.py, config.py

``utils/vector_store.py`` (the FAISS wrapper this repository replaces) is deliberately NOT staged:
the harness injects the drop-in class under that module name.  A manifest with the SHA-256 of every
staged file is written next to them; ``tests/golden/reference_suite.sha256`` (committed: digests, not
source) pins it, so a test run can prove that what ran on the GPU box is the unmodified reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
from typing import Dict, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "oracle", "_ref", "reference_suite")
PINNED = os.path.join(ROOT, "tests", "golden", "reference_suite.sha256")
DEFAULT_SOURCE = "/root/reference"

FILES = [
    "config.py",
    "core/__init__.py",
    "core/searcher.py",
    "tests/__init__.py",
    "tests/helpers.py",
    "tests/test_searcher.py",
    "tests/test_vector_store.py",
    "utils/__init__.py",
    "utils/path_utils.py",
    "utils/structured_analysis.py",
]


def _digest(path: str) -> str:
    with open(path, "rb") as handle:
        return hashlib.sha256(handle.read()).hexdigest()


def manifest_of(folder: str) -> Dict[str, str]:
    return {rel: _digest(os.path.join(folder, rel)) for rel in FILES if os.path.exists(os.path.join(folder, rel))}


def stage(source: str = DEFAULT_SOURCE, write_pin: bool = False) -> Optional[str]:
    """Copy the files from ``source`` when it exists; returns the staged folder (or None when there is
    neither a reference checkout nor an earlier staging)."""
    if not os.path.isdir(os.path.join(source, "tests")):
        return STAGED if os.path.isdir(os.path.join(STAGED, "tests")) else None
    for rel in FILES:
        dst = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source, rel), dst)
    manifest = manifest_of(STAGED)
    with open(os.path.join(STAGED, "MANIFEST.json"), "w", encoding="utf-8") as handle:
        json.dump(manifest, handle, indent=1, sort_keys=True)
    if write_pin:
        with open(PINNED, "w", encoding="utf-8") as handle:
            for rel in FILES:
                handle.write(f"{manifest[rel]}  {rel}\n")
    return STAGED


def pinned_manifest() -> Dict[str, str]:
    out: Dict[str, str] = {}
    with open(PINNED, "r", encoding="utf-8") as handle:
        for line in handle:
            digest, rel = line.split()
            out[rel] = digest
    return out


def locate() -> Optional[str]:
    """Where the reference's suites can be run from: the checkout itself, else the staged copy."""
    env = os.environ.get("PSX_REFERENCE")
    for cand in (env, DEFAULT_SOURCE, STAGED):
        if cand and os.path.isdir(os.path.join(cand, "tests")) and os.path.isdir(os.path.join(cand, "core")):
            return cand
    return None


if __name__ == "__main__":
    print(stage(write_pin=True))
