"""Per-phase cycle breakdown of the intra wavefront kernel (needs a build with
MPTC_PHASE_TIMING=1 python -m mptc_b200.build --force)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, SA, THR = 1920, 1080, 16, 50
frames = np.stack([make_frame(W, H, f) for f in (0, 15, 30, 45)])
ctx = capi.Context(0)
ctx.set_schedule(1, 0, 0)   # one lane: the four intra frames share one wavefront launch
ctx.seq_reserve(W, H, 4)
ctx.seq_upload(frames)
L = capi.load()
buf = (C.c_ulonglong * 16)()
for it in range(2):
    L.mptc_debug_phase_cycles(buf, 1)
    ctx.seq_encode(0, 4, SA, THR, 1)   # gop 1: four intra frames in one launch
    ctx.sync()
L.mptc_debug_phase_cycles(buf, 0)
names = ["loop/todo", "wait rows above", "load+hash+ids", "remap+wordinfo", "evaluate", "rows above scan", "endpoint refit+write+sync", "in-row decisions", "final release"]
tot = sum(buf[i] for i in range(9))
groups = buf[11]
print("intra ms", ctx.last_encode_ms("intra"), "groups", groups, "avg distinct words/group", buf[10] / max(groups, 1))
for i, n in enumerate(names):
    print(f"{n:24s} {buf[i] / max(groups,1):10.0f} cycles/group  {100.0 * buf[i] / tot:5.1f}%")
groups = max(groups, 1)
print(f"decider detail per group: near_update calls {buf[15] / groups:.2f}, cycles in near_update {buf[13] / groups:.0f} "
      f"(of which polling {buf[14] / groups:.0f}), words added late {buf[12] / groups:.2f}")
