#!/usr/bin/env python
"""SASS census of the shipped libpsx.so: how often the Blackwell-only instructions occur (cuobjdump -sass).
`python tools/sass_census.py > profiles/r2_sass_census.txt`; tests/test_abi.py asserts the same counts are non-zero."""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = {
    "UTCHMMA": "tcgen05.mma (5th-gen tensor core, operands from shared memory, accumulator in TMEM)",
    "UTCHMMA.2CTA": "tcgen05.mma.cta_group::2 (one MMA over the two SMs of a CTA pair)",
    "LDTM": "tcgen05.ld (TMEM -> registers, the epilogue's read-back)",
    "UTMALDG": "cp.async.bulk.tensor (TMA tiled load through a tensor map)",
    "UBLKCP": "cp.async.bulk (TMA 1-D bulk copy: the scan's row stream)",
    "SYNCS": "mbarrier operations",
    "UTCBAR": "tcgen05.commit (MMA completion -> mbarrier)",
    "UTCATOMSWS": "tcgen05.alloc / dealloc (TMEM allocation)",
    "HMMA": "legacy mma.sync tensor-core path (must stay 0)",
}


def census(lib: str):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.Counter()
    per_kernel = collections.defaultdict(collections.Counter)
    kernel = "?"
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kernel = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for key in MNEMONICS:
            if op == key or op.startswith(key + ".") or (key == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                if key == "UTCHMMA" and ".2CTA" in op:
                    continue
                counts[key] += 1
                per_kernel[kernel][key] += 1
    return counts, per_kernel, "sm_100a" in out


def main():
    sys.path.insert(0, ROOT)
    from photo_search_engine_b200 import _native

    counts, per_kernel, is_100a = census(_native.LIB_PATH)
    print(f"library: {os.path.relpath(_native.LIB_PATH, ROOT)}   arch sm_100a: {is_100a}")
    for key, what in MNEMONICS.items():
        print(f"{counts.get(key, 0):6d}  {key:14s} {what}")
    print("\nper kernel (demangled prefix):")
    for kernel, c in sorted(per_kernel.items()):
        try:
            name = subprocess.run(["c++filt", kernel], capture_output=True, text=True).stdout.strip()
        except Exception:
            name = kernel
        print(f"  {name[:110]}")
        print("      " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
