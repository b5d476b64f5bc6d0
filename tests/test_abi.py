"""The C-ABI shared library: loads, exports exactly what include/psx.h declares, and refuses
to compute without a B200 (no CPU fallback).  No kernel is launched here."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import pytest

from tests.conftest import ROOT, has_gpu

HEADER = os.path.join(ROOT, "include", "psx.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psx_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from photo_search_engine_b200 import _native

    lib = _native.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in psx.h but not exported by libpsx.so"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared  # the ctypes table covers the whole header
    assert lib.psx_abi_version() == 2
    assert _native.kpad(100) == 128 and _native.kpad(1) == 32 and _native.kpad(2048) == 2048


def test_library_is_sm100a_with_tma():
    """The shipped SASS is sm_100a and the scan really uses the bulk async copy engine."""
    from photo_search_engine_b200 import _native

    out = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "UBLKCP" in out.stdout and "SYNCS" in out.stdout


def test_sass_census_shows_tcgen05_tmem_tma():
    """The batched path is genuinely tcgen05 / TMEM / tensor-map TMA and nothing falls back to mma.sync
    (census committed as profiles/r2_sass_census.txt, tool: tools/sass_census.py)."""
    import importlib.util

    from photo_search_engine_b200 import _native

    spec = importlib.util.spec_from_file_location("sass_census", os.path.join(ROOT, "tools", "sass_census.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        counts, per_kernel, is_100a = mod.census(_native.LIB_PATH)
    except Exception:
        pytest.skip("cuobjdump unavailable")
    assert is_100a
    assert counts["UTCHMMA"] >= 8 and counts["UTCHMMA.2CTA"] >= 4     # cta_group::1 and cta_group::2 MMAs
    assert counts["LDTM"] >= 4 and counts["UTMALDG"] >= 4             # TMEM read-back, tiled TMA loads
    assert counts["UBLKCP"] >= 100 and counts["UTCBAR"] >= 4          # bulk row stream, MMA -> mbarrier commits
    assert counts["HMMA"] == 0                                          # no legacy tensor-core path anywhere
    gemm = [c for name, c in per_kernel.items() if "gemm_filter" in name]
    assert gemm and all(c["UTCHMMA"] + c["UTCHMMA.2CTA"] > 0 and c["LDTM"] > 0 and c["UTMALDG"] > 0 for c in gemm)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a machine without a GPU")
def test_fails_loudly_without_gpu(tmp_path):
    from photo_search_engine_b200 import _native
    from photo_search_engine_b200.vector_store import VectorStore

    with pytest.raises(RuntimeError) as err:
        _native.NativeIndex(8)
    assert "no CPU fallback" in str(err.value)
    with pytest.raises(RuntimeError):
        VectorStore(8, str(tmp_path / "i"), str(tmp_path / "m"))
    # argument validation happens before any device is touched
    with pytest.raises(ValueError):
        _native.NativeIndex(0)
    lib = _native.load_library()
    assert lib.psx_ntotal(None) == 0 and lib.psx_destroy(None) == 0


def test_missing_library_is_an_import_error(tmp_path):
    code = "import photo_search_engine_b200.vector_store"
    env = dict(os.environ, PSX_LIB=str(tmp_path / "nope.so"), PYTHONPATH=ROOT)
    proc = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert proc.returncode != 0 and "ImportError" in proc.stderr
