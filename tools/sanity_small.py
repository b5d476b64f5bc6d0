#!/usr/bin/env python
"""Small end-to-end exercise of every kernel (used under compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from photo_search_engine_b200 import _native as N

rng = np.random.default_rng(0)
for d, n, dt in [(8, 300, 0), (100, 900, 0), (768, 700, 1), (1024, 1500, 0), (4100, 200, 0), (64, 70000, 2)]:
    x = rng.standard_normal((n, d)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((3, d)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    ix = N.NativeIndex(d, 0, dt, 0); ix.add(x)
    ix.set_attrs(0, (np.arange(n, dtype=np.uint64) + np.uint64(1)))
    flt = N.PsxFilter(flags=N.F_NEED_DT | N.F_START | N.F_END, start=n // 4, end=n // 2)
    for k in (1, 50, 300):
        D, I = ix.search(q, k); Df, If = ix.search(q, k, flt)
        assert (If[If >= 0] >= n // 4 - 1).all()
    ix.reconstruct(3); ix.read_rows(0, 10)
    if n >= 65536:
        ix.set_tunable("batch_min", 2); qb = rng.standard_normal((9, d)).astype(np.float32); ix.search(qb, 10)
    ix.close()
print("sanity ok")
