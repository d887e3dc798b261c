"""BASELINE.json configs[3]: 1080p search-window / error-threshold sweep of the index-reuse kernels.
One GOP of 15 frames (1 intra + 14 inter), sa in {2,4,8,16,32} x thr in {0,10,50,200}; prints
Mpixel/s (device time, frames resident), the share of blocks found by the inter / intra search,
unique blocks, and candidate positions per second (SURVEY.md 8d's unit).
usage: python profiles/window_sweep.py [W H FRAMES]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, FRAMES = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 15)
GOP = 15
nb = (W // 4) * (H // 4)
pin = capi.PinnedArray((FRAMES, H, W, 3), np.uint8)
for f in range(FRAMES):
    pin.array[f] = make_frame(W, H, f)
ctx = capi.Context(0)
ctx.seq_reserve(W, H, FRAMES)
ctx.seq_upload(pin.array)
ctx.sync()
print(f"# {W}x{H} x{FRAMES} frames, gop {GOP}; device time of mptc_gpu_seq_encode, best of 3")
print("sa thr   ms/seq  Mpixel/s  inter% intra% unique%  Gcand/s  stage ms: fit inter intra")
for sa in (2, 4, 8, 16, 32):
    for thr in (0, 10, 50, 200):
        ms = []
        for _ in range(4):
            ctx.seq_encode(0, FRAMES, sa, thr, GOP)
            ms.append(ctx.last_encode_ms("total"))
        t = min(ms[1:])
        out = ctx.seq_download(0, FRAMES, want=("motion", "unique"))
        m = out["motion"].reshape(FRAMES, nb, 2)
        uniq = (m[:, :, 0] == 255) & (m[:, :, 1] == 255)
        inter = ((m[:, :, 0] & 0x80) != 0) & ((m[:, :, 1] & 0x80) != 0) & ~uniq
        intra = ~uniq & ~inter
        assert int(uniq.sum()) == int(out["n_unique"].sum())
        ci, ca = ctx.last_candidate_count()
        tot = FRAMES * nb
        print(f"{sa:2d} {thr:3d} {t:8.2f} {FRAMES * W * H / t / 1e3:9.0f}  {100 * inter.sum() / tot:5.1f} {100 * intra.sum() / tot:6.1f} "
              f"{100 * uniq.sum() / tot:7.2f} {(ci + ca) / t / 1e6:8.1f}   {ctx.last_encode_ms('fit'):.2f} "
              f"{ctx.last_encode_ms('inter'):.2f} {ctx.last_encode_ms('intra'):.2f}", flush=True)
