"""Per-phase cycle breakdown of the second-generation intra wavefront (k_intra_rows); needs a build
with MPTC_PHASE_TIMING=1 python -m mptc_b200.build --force.  Four 1080p intra frames in one launch."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, SA, THR = 1920, 1080, 16, int(os.environ.get("THR", "50"))
NF = int(os.environ.get("NF", "4"))
frames = np.stack([make_frame(W, H, f) for f in (0, 15, 30, 45)[:NF]])
ctx = capi.Context(0)
ctx.set_schedule(1, int(os.environ.get("WAVE_ROWS", "0")), 0)   # one lane: the intra frames share one wavefront launch
ctx.seq_reserve(W, H, NF)
ctx.seq_upload(frames)
L = capi.load()
buf = (C.c_ulonglong * 24)()
TIMED = hasattr(L, "mptc_debug_rows_cycles")
ms = []
for it in range(5):
    if TIMED:
        L.mptc_debug_rows_cycles(buf, 1)
    ctx.seq_encode(0, NF, SA, THR, 1)   # gop 1: all frames intra, one launch
    ctx.sync()
    ms.append(round(ctx.last_encode_ms("intra"), 3))
if not TIMED:
    print("intra ms per launch (no MPTC_PHASE_TIMING build):", ms, "split", os.environ.get("MPTC_ROW_SPLIT", "default"))
    sys.exit(0)
L.mptc_debug_rows_cycles(buf, 0)
LIGHT = buf[11] == 0   # an MPTC_PHASE_TIMING=2 build: only the trace
if LIGHT:
    print("intra ms per launch (trace-only build):", ms)
g = max(buf[11], 1)
print("intra ms", round(ctx.last_encode_ms("intra"), 3), "fast-path groups", buf[11], "avg distinct words/group", round(buf[10] / g, 1))
names = ["loop/todo", "A: far rows + snapshot + clear", "B: window load + hash", "C: ids + constants + evaluation", "C: rows above scan",
         "D: own row (barrier to barrier)"]
tot = max(sum(buf[i] for i in range(6)), 1)
for i, nm in enumerate(names):
    print(f"{nm:34s} {buf[i] / g:10.0f} cycles/group  {100.0 * buf[i] / tot:5.1f}%")
print(f"decider per group: loop {buf[8] / g:.0f} cycles, of which waiting for the partner's left part {buf[6] / g:.0f}, "
      f"waiting for near-row partials {buf[7] / g:.0f} in {buf[9] / g:.1f} waits; words added late {buf[12] / g:.2f}")
print(f"merger (row above) per group: {buf[14] / g:.1f} polls, {buf[13] / g:.1f} with news, {buf[15] / max(buf[13], 1):.1f} words per batch")

if os.environ.get("TRACE"):
    bh, ng = H // 4, (W // 4 + 31) // 32
    tr = (C.c_ulonglong * (512 * 16 * 6))()
    L.mptc_debug_rows_trace(tr, 512 * 16 * 6)
    a = np.frombuffer(tr, dtype=np.uint64).reshape(512, 16, 6)[:bh, :ng].astype(np.float64)
    t0 = a[a > 0].min()
    a = (a - t0) / 1e3   # us
    np.set_printoptions(linewidth=220, precision=1, suppress=True)
    print("times in us since the first event; columns: wait-far start, load start, own-row start, first decision, last decision")
    for by in (0, 1, 2, 3, 4, 5, 6, 8, 16, 32, 64, 100, 101, 102, 103, 104, 200, 268, 269):
        for gi in (0, 1, 7, 14):
            print(f"row {by:3d} group {gi:2d}: ", a[by, gi, :5])
    last = a[:, ng - 1, 4]
    print("row completion times (us), every 10th row:", last[::10])
    d = np.diff(last)
    print(f"lag between consecutive rows' completion: mean {d.mean():.2f} us, median {np.median(d):.2f}, first 40 rows mean {d[:40].mean():.2f}, rows 100-200 mean {d[100:200].mean():.2f}")
    g0 = a[:, 0, 3]
    print(f"lag between consecutive rows' FIRST decision: mean {np.diff(g0).mean():.2f} us")
    per_group = (a[:, 1:, 4] - a[:, :-1, 4])
    print(f"time between last decisions of consecutive groups of a row: mean {per_group.mean():.2f} us, rows 100-200 {per_group[100:200].mean():.2f}")
    print(f"decision span inside a group (first to last decision): mean {(a[:, :, 4] - a[:, :, 3]).mean():.2f} us")
    print(f"build (load start to own-row start): mean {(a[:, :, 2] - a[:, :, 1]).mean():.2f} us; far wait mean {(a[:, :, 1] - a[:, :, 0]).mean():.2f} us; "
          f"own-row start to first decision mean {(a[:, :, 3] - a[:, :, 2]).mean():.2f} us")
    if LIGHT:
        sys.exit(0)
    st = (C.c_ulonglong * (8 * 512))()
    L.mptc_debug_rows_steps(st, 8 * 512)
    s = (np.frombuffer(st, dtype=np.uint64).reshape(8, 512)[:, : W // 4].astype(np.float64) - t0) / 1e3
    base = s[0, 192]
    print("decision times (us, relative to row 100 block 192) of blocks 192..287 (groups 6-8), rows 100..103:")
    for r in range(4):
        print(f"row {100 + r}:", np.array2string(s[r, 192:288] - base, precision=1, max_line_width=230))
    print("for each block of row 101: time since row 100 decided the block 16 to its right (the last one its window needs):")
    print(np.array2string(s[1, 192:272] - s[0, 208:288], precision=1, max_line_width=230))
