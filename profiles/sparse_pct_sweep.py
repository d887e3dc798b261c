"""Which kernel should take the blocks an inter frame's inter search leaves over?  K3s (sparse,
direct evaluation per item) up to MPTC_SPARSE_MAX_PCT percent of the frame's blocks, the row
wavefront K3 above that.  Sweeps the threshold on the leftover-heavy corner of BASELINE.json
configs[3] (err_threshold 0) and on the headline parameters.
usage: python profiles/sparse_pct_sweep.py"""
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
print(f"# {W}x{H} x{FRAMES} frames, gop {GOP}; device ms of mptc_gpu_seq_encode (best of 2), stage ms: inter intra")
print("sa thr  pct   ms/seq  inter  intra")
ref = {}
for sa, thr in ((16, 0), (8, 0), (2, 0), (16, 50), (2, 50)):
    for pct in (0, 1, 3, 10, 50):
        os.environ["MPTC_SPARSE_MAX_PCT"] = str(pct)
        ctx = capi.Context(0)
        ctx.seq_reserve(W, H, FRAMES)
        ctx.seq_upload(pin.array)
        best = None
        for _ in range(3):
            ctx.seq_encode(0, FRAMES, sa, thr, GOP)
            t = (ctx.last_encode_ms("total"), ctx.last_encode_ms("inter"), ctx.last_encode_ms("intra"))
            best = t if best is None or t[0] < best[0] else best
        out = ctx.seq_download(0, FRAMES, want=("blocks",))["blocks"]
        key = (sa, thr)
        if key in ref:
            assert np.array_equal(ref[key], out), "results depend on the leftover kernel"
        ref[key] = out
        print(f"{sa:2d} {thr:3d}  {pct:3d}  {best[0]:7.2f} {best[1]:6.2f} {best[2]:6.2f}", flush=True)
        ctx.close()
