import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from mptc_b200 import capi
from mptc_b200.synth import make_frame
W,H,N,SA,THR,GOP=1920,1080,30,16,50,15
frames=np.stack([make_frame(W,H,f) for f in range(N)])
ctx=capi.Context(0); ctx.seq_reserve(W,H,N); ctx.seq_upload(frames); ctx.seq_encode(0,N,SA,THR,GOP); ctx.sync()
ms=[]
for _ in range(6):
    ctx.seq_decode(0,N,SA,GOP,rgb=True); ms.append((ctx.last_decode_ms("rgb"), ctx.last_decode_ms("words"), ctx.last_decode_ms("planes")))
print(min(m[0] for m in ms), "GB/s", N*(W//4)*(H//4)*56/min(m[0] for m in ms)/1e6, "words", min(m[1] for m in ms), "planes", min(m[2] for m in ms))
