"""Debug helper: first mismatching blocks of a small sequence against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mptc_b200 import capi
from mptc_b200.synth import make_sequence
from oracle import port

w, h, n, sa, thr, gop = [int(x) for x in (sys.argv[1:7] if len(sys.argv) > 6 else (320, 192, 6, 4, 20, 3))]
frames = make_sequence(w, h, n)
ctx = capi.Context(0)
out = ctx.encode_sequence(frames, sa, thr, gop)
prev = None
bw = w // 4
for i in range(n):
    init = port.dxt1_fit(frames[i])
    blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
    prev = blocks
    bad = np.nonzero(out["blocks"][i] != blocks)[0]
    m = out["motion"][i].reshape(-1, 2); mr = motion.reshape(-1, 2)
    badm = np.nonzero((m != mr).any(axis=1))[0]
    print(f"frame {i} intra={i % gop == 0}: {bad.size} blocks differ, {badm.size} motion differ, unique {out['n_unique'][i]} vs {unique.size}")
    for b in badm[:6]:
        print(f"   block {b} (x={b % bw}, y={b // bw}): motion got {tuple(m[b])} want {tuple(mr[b])}; word got {int(out['blocks'][i][b]) >> 32:#010x} want {int(blocks[b]) >> 32:#010x} init {int(init[b]) >> 32:#010x}")
