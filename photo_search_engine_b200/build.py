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
OBJ_DIR = os.path.join(HERE, "build")
HEADERS = [os.path.join("..", "..", "include", "psx.h")]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu to an object (in parallel: the scan kernel is instantiated per stored
    type and metric in its own translation unit) and link them into libpsx.so."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    extra = ["-DPSX_DEBUG_KERNELS"] if os.environ.get("PSX_DEBUG_KERNELS") == "1" else []
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", obj, src]
        return cmd, obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, srcs))
    log = []
    for cmd, _, proc in results:
        log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
    failed = [proc for _, _, proc in results if proc.returncode != 0]
    if failed:
        raise RuntimeError("nvcc failed:\n" + failed[0].stderr[-4000:])
    cmd = [nvcc, *LINK_FLAGS, "-o", LIB + ".tmp", *[obj for _, obj, _ in results]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log.append(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stderr[-4000:])
    os.replace(LIB + ".tmp", LIB)
    with open(os.path.join(HERE, "libpsx.build.log"), "w") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
