"""In-tree build of the native library (``libpsx.so``) for sm_100a.

``python -m photo_search_engine_b200.build`` or ``build_native()``.  nvcc cross-compiles
without a GPU; the resulting ``.so`` is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsx.so")
SOURCES = ["psx_api.cu"]
HEADERS = ["psx_common.cuh", "psx_scan.cuh", "psx_aux.cuh", os.path.join("..", "..", "include", "psx.h")]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-shared",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    extra = ["-DPSX_DEBUG_KERNELS"] if os.environ.get("PSX_DEBUG_KERNELS") == "1" else []
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB + ".tmp", *srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stderr[-4000:])
    os.replace(LIB + ".tmp", LIB)
    with open(os.path.join(HERE, "libpsx.build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
