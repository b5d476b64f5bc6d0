"""Test configuration.

* ``-m "not gpu"``: oracle vs golden fixtures, host logic of the drop-in class (driven through an
  oracle-backed fake backend), C-ABI surface (library loads, exports every symbol of
  include/psx.h, fails loudly without a GPU).
* ``-m gpu``: parity tests proper -- CUDA path through the C ABI vs the oracle.
"""
from __future__ import annotations

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _locate_reference() -> str:
    """The reference checkout (build container), else the unmodified copy of its hot-path suites that
    ``__graft_entry__.build()`` staged under oracle/_ref/ (git-ignored, travels to the GPU box)."""
    from oracle.stage_reference import DEFAULT_SOURCE, locate

    return locate() or DEFAULT_SOURCE


REFERENCE = _locate_reference()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    from photo_search_engine_b200.build import build_native

    build_native()


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    _ensure_built()
    yield


@pytest.fixture()
def fake_backend(monkeypatch):
    """Swap the GPU backend of ``VectorStore`` for the oracle-backed fake (host-logic tests)."""
    from photo_search_engine_b200.vector_store import VectorStore
    from tests._fake_backend import FakeIndex

    monkeypatch.setattr(VectorStore, "_index_factory", staticmethod(FakeIndex))
    return FakeIndex


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
