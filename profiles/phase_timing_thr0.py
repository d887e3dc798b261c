"""Phase breakdown of the row wavefront K3 on the leftover-heavy corner (err_threshold 0): an
intra frame alone, then intra + inter frame with K3s switched off (MPTC_SPARSE_MAX_PCT=0) so that
the wavefront takes the inter frame's leftovers.  Needs MPTC_PHASE_TIMING=1 python -m mptc_b200.build --force."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["MPTC_SPARSE_MAX_PCT"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, SA = 1920, 1080, 16
THR = int(sys.argv[1]) if len(sys.argv) > 1 else 0
frames = np.stack([make_frame(W, H, f) for f in (0, 1)])
ctx = capi.Context(0)
ctx.set_schedule(1, 0, 0)
ctx.seq_reserve(W, H, 2)
ctx.seq_upload(frames)
L = capi.load()
buf = (C.c_ulonglong * 16)()
names = ["loop/todo", "wait rows above", "load+hash+ids", "remap+wordinfo", "evaluate", "rows above scan", "endpoint refit+write+sync", "in-row decisions", "final release"]
prev = None
for label, n in (("intra frame only", 1), ("intra + inter frame", 2)):
    for it in range(2):
        L.mptc_debug_phase_cycles(buf, 1)
        ctx.seq_encode(0, n, SA, THR, 2)
        ctx.sync()
    L.mptc_debug_phase_cycles(buf, 0)
    cur = [int(buf[i]) for i in range(16)]
    show = cur if prev is None else [a - b for a, b in zip(cur, prev)]
    what = label if prev is None else "inter frame (difference)"
    groups = max(show[11], 1)
    tot = max(sum(show[:9]), 1)
    print(f"== {what}: thr {THR}, intra stage {ctx.last_encode_ms('intra'):.2f} ms (cumulative), groups {groups}, "
          f"avg distinct words/group {show[10] / groups:.1f}")
    for i, nm in enumerate(names):
        print(f"  {nm:26s} {show[i] / groups:10.0f} cycles/group  {100.0 * show[i] / tot:5.1f}%")
    prev = cur
out = ctx.seq_download(0, 2, want=("motion",))
m = out["motion"][1].reshape(-1, 2)
uniq = (m[:, 0] == 255) & (m[:, 1] == 255)
inter = ((m[:, 0] & 0x80) != 0) & ((m[:, 1] & 0x80) != 0) & ~uniq
print(f"inter frame: {100 * inter.mean():.1f}% inter, {100 * (~inter & ~uniq).mean():.1f}% intra, {100 * uniq.mean():.1f}% unique")
