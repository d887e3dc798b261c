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


def read_trace():
    cap = 1 << 16
    tr = (C.c_ulonglong * (3 * cap))()
    n = L.mptc_debug_k2w_trace(tr, cap, 1)
    a = np.frombuffer(tr, dtype=np.uint64)[:3 * n].reshape(n, 3)
    return a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), (a[:, 2] & 0xFFFF).astype(int), ((a[:, 2] >> 16) & 0xFFFF).astype(int)


def occupancy(t0, t1, sm, title, bin_us=20.0):
    """Running K2 CTAs over time (2 x 148 slots)."""
    lo, hi = t0.min(), t1.max()
    span = (hi - lo) / 1e3
    slots = 2 * 148
    busy = ((t1 - t0).sum() / 1e3) / (slots * span)
    print(f"{title}: {len(t0)} CTAs, span {span / 1e3:.3f} ms, CTA-slot utilisation {busy:.2f}")
    nb = int(span / bin_us) + 1
    occ = np.zeros(nb)
    for a, b in zip((t0 - lo) / 1e3, (t1 - lo) / 1e3):
        i0, i1 = int(a / bin_us), int(b / bin_us)
        if i0 == i1:
            occ[i0] += (b - a) / bin_us
        else:
            occ[i0] += (i0 + 1) - a / bin_us
            occ[i0 + 1:i1] += 1
            occ[i1] += b / bin_us - i1
    hist = np.histogram(occ, bins=[0, 1, 74, 148, 222, 280, 297])[0]
    print("  time share by running CTAs: " + ", ".join(f"{n}: {100.0 * h / nb:.0f}%" for n, h in zip(("0", "1-73", "74-147", "148-221", "222-279", "280-296"), hist)))
    return occ


if wide and hasattr(L, "mptc_debug_k2w_trace"):
    L.mptc_debug_k2w_trace.restype = C.c_int
    read_trace()
    ctx.seq_encode(0, N, SA, THR, GOP)
    ctx.sync()
    t0, t1, sm, words = read_trace()
    occupancy(t0, t1, sm, "one lane, one GOP")
    dur = (t1 - t0) / 1e3
    print(f"  CTA duration mean {dur.mean():.1f} / median {np.median(dur):.1f} / max {dur.max():.1f} us; distinct words per tile mean {words.mean():.0f} max {words.max()}, over the table {(words > 224).sum()}")
    # four lanes, four GOPs: the benchmark's schedule
    N4 = 60
    frames4 = np.stack([make_frame(W, H, f) for f in range(N4)])
    ctx4 = capi.Context(0)
    ctx4.seq_reserve(W, H, N4)
    ctx4.seq_upload(frames4)
    for _ in range(2):
        ctx4.seq_encode(0, N4, SA, THR, GOP)
    ctx4.sync()
    read_trace()
    ctx4.seq_encode(0, N4, SA, THR, GOP)
    ctx4.sync()
    print("four lanes: step", round(ctx4.last_encode_ms("total"), 3), "ms")
    t0, t1, sm, words = read_trace()
    occ = occupancy(t0, t1, sm, "four lanes, four GOPs")
    print("  running CTAs per 20 us bin, first 3 ms: " + " ".join(f"{int(round(x))}" for x in occ[:150]))
