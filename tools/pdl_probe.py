import numpy as np, sys, time
sys.path.insert(0, '/root/repo')
from photo_search_engine_b200 import _native as N
rng = np.random.default_rng(1)
def unit(n,d):
    x = rng.standard_normal((n,d)).astype(np.float32); return x/np.linalg.norm(x,axis=1,keepdims=True)
for (n,d,nq,k) in [(131072,4096,9,20),(150000,1024,200,100),(110000,64,40,512)]:
    x = unit(n,d); q = unit(nq,d)
    ix = N.NativeIndex(d); ix.add(x)
    ix.set_tunable("batch_min",0); Ds,Is = ix.search(q,k)
    for pdl in (0,1):
        ix.set_tunable("batch_min",2); ix.set_tunable("batch_pdl",pdl)
        b0 = ix.batch_stats()
        for rep in range(3):
            Db,Ib = ix.search(q,k)
            print((n,d,nq,k), "pdl",pdl,"rep",rep,"equal",np.array_equal(Ib,Is) and np.array_equal(Db,Ds), "stats", tuple(a-b for a,b in zip(ix.batch_stats(),b0)), flush=True)
    ix.close()
