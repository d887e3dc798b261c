import sys,time; import os; sys.path.insert(0, os.getcwd())
import numpy as np
from mptc_b200 import capi
rng=np.random.default_rng(0)
nb=129600; ps=512*320
motion=np.clip(rng.normal(144,2,2*nb),0,255).astype(np.uint8)
planes=np.clip(rng.normal(128,3,6*ps),0,255).astype(np.uint8)
nsym=2*nb+6*ps
best_p=best_s=best_d=0
sym=np.clip(rng.normal(128,3,2_000_000),0,255).astype(np.uint8)
code=capi.arith_encode(sym)
for rep in range(4):
    t=time.time()
    for _ in range(3): capi.frame_payload(motion,planes,10,1)
    best_p=max(best_p,nsym/((time.time()-t)/3)/1e6)
    t=time.time()
    for _ in range(3):
        capi.arith_encode(motion); capi.arith_encode(planes[:ps]); capi.arith_encode(planes[ps:3*ps]); capi.arith_encode(planes[3*ps:4*ps]); capi.arith_encode(planes[4*ps:])
    best_s=max(best_s,nsym/((time.time()-t)/3)/1e6)
    t=time.time(); capi.arith_decode(code,sym.size); best_d=max(best_d,sym.size/(time.time()-t)/1e6)
print("pairs %.1f  singles %.1f  decode %.1f Msym/s"%(best_p,best_s,best_d))
