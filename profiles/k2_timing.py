"""Per-phase cycle breakdown of K2 (k_inter_search_wide, or k_inter_search_tiled with MPTC_K2=tiled); needs a build with
MPTC_EXTRA_NVCC_FLAGS=-DMPTC_K2_PHASE_TIMING python -m mptc_b200.build --force.  One GOP of 15 1080p frames."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 15, 16, int(os.environ.get("THR", "50")), 15
frames = np.stack([make_frame(W, H, f) for f in range(N)])
ctx = capi.Context(0)
ctx.set_schedule(1, 0, 0)
ctx.seq_reserve(W, H, N)
ctx.seq_upload(frames)
L = capi.load()
buf = (C.c_ulonglong * 12)()
wide = os.environ.get("MPTC_K2", "wide") != "tiled"
read = L.mptc_debug_k2w_cycles if wide else L.mptc_debug_k2_cycles
for it in range(3):
    read(buf, 1)
    ctx.seq_encode(0, N, SA, THR, GOP)
    ctx.sync()
read(buf, 0)
tiles = ctx.last_work_count()["inter_tiles"]
names = (["clear table", "window load + hash", "ids", "per-word constants + marks", "evaluation (warp 0)", "window scan (warp 0)", "resolve + apply", "-"] if wide else
         ["clear table", "window load + hash", "ids", "per-word constants", "evaluation", "window scan", "resolve", "apply"])
tot = sum(buf[i] for i in range(8))
print("inter ms", round(ctx.last_encode_ms("inter"), 3), "tiles", tiles, "cycles per tile (thread 0's view, two CTAs share an SM):", round(tot / tiles))
for i, nm in enumerate(names):
    print(f"{nm:22s} {buf[i] / tiles:9.0f} cycles/tile  {100.0 * buf[i] / tot:5.1f}%")
