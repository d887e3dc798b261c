"""CTAs per frame of the leftover kernel K3s (MPTC_SPARSE_CTAS) on leftover-heavy frames
(err_threshold 0: 10-45 % of an inter frame's blocks go to the intra search).
usage: python profiles/sparse_ctas_sweep.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, FRAMES, GOP = 1920, 1080, 15, 15
pin = capi.PinnedArray((FRAMES, H, W, 3), np.uint8)
for f in range(FRAMES):
    pin.array[f] = make_frame(W, H, f)
print(f"# {W}x{H} x{FRAMES} frames, gop {GOP}; device ms of mptc_gpu_seq_encode (best of 3), stage ms: inter intra")
print("sa thr  ctas   ms/seq  inter  intra")
for sa, thr in ((16, 0), (8, 0), (2, 0), (16, 50)):
    for ctas in (148, 296, 592, 1184):
        os.environ["MPTC_SPARSE_CTAS"] = str(ctas)
        os.environ["MPTC_SPARSE_MAX_PCT"] = "100"
        ctx = capi.Context(0)
        ctx.seq_reserve(W, H, FRAMES)
        ctx.seq_upload(pin.array)
        best = None
        for _ in range(3):
            ctx.seq_encode(0, FRAMES, sa, thr, GOP)
            t = (ctx.last_encode_ms("total"), ctx.last_encode_ms("inter"), ctx.last_encode_ms("intra"))
            best = t if best is None or t[0] < best[0] else best
        print(f"{sa:2d} {thr:3d}  {ctas:4d}  {best[0]:7.2f} {best[1]:6.2f} {best[2]:6.2f}", flush=True)
        ctx.close()
