import sys, numpy as np
sys.path.insert(0, '/root/repo')
from photo_search_engine_b200 import _native as N
rng = np.random.default_rng(0)
n, d, nq, k = 70000, 64, 5, 10
x = rng.standard_normal((n, d)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
q = rng.standard_normal((nq, d)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
ix = N.NativeIndex(d); ix.add(x)
ix.set_tunable("batch_min", 0); Ds, Is = ix.search(q, k)
ix.set_tunable("batch_min", 2)
try:
    Db, Ib = ix.search(q, k)
    print("ids equal", np.array_equal(Ib, Is), "scores equal", np.array_equal(Db, Ds), ix.batch_stats())
    print(Ib[:2], Is[:2])
except Exception as e:
    print("ERR", e)
