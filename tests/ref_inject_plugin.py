"""pytest plugin used by tests/test_reference_suite.py: makes the reference's UNMODIFIED
``tests/test_vector_store.py`` / ``tests/test_searcher.py`` import this repository's drop-in class
as ``utils.vector_store.VectorStore`` (what INTEGRATION.md asks a maintainer to do with a
one-line re-export).

PSX_REF_BACKEND=fake  -> host logic over the oracle-backed fake (CPU container)
PSX_REF_BACKEND=gpu   -> the real CUDA backend
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from photo_search_engine_b200.vector_store import VectorStore  # noqa: E402

if os.environ.get("PSX_REF_BACKEND", "fake") == "fake":
    # loaded by path under a private name: the name ``tests`` must stay free for the reference's
    # own ``tests`` package (its files do ``from tests.helpers import ...``)
    import importlib.util

    _spec = importlib.util.spec_from_file_location("psx_fake_backend", os.path.join(ROOT, "tests", "_fake_backend.py"))
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    VectorStore._index_factory = staticmethod(_mod.FakeIndex)

shim = types.ModuleType("utils.vector_store")
shim.VectorStore = VectorStore
shim.__file__ = __file__
sys.modules["utils.vector_store"] = shim
