"""Sweeps the scheduling knobs (GOP lanes, wavefront CTAs per frame) on the bench workload and
checks that every setting produces identical results.
usage: python profiles/sched_sweep.py [W H FRAMES]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, FRAMES = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 60)
SA, THR, GOP = 16, 50, 15
nb = (W // 4) * (H // 4)
pbw, pbh = (W // 4 + 63) // 64 * 64, (H // 4 + 63) // 64 * 64
pin = capi.PinnedArray((FRAMES, H, W, 3), np.uint8)
for f in range(FRAMES):
    pin.array[f] = make_frame(W, H, f)
pins = {"blocks": capi.PinnedArray((FRAMES, nb), np.uint64), "motion": capi.PinnedArray((FRAMES, 2 * nb), np.uint8),
        "unique": capi.PinnedArray((FRAMES, nb), np.uint32), "n_unique": capi.PinnedArray((FRAMES,), np.uint32),
        "planes": capi.PinnedArray((FRAMES, 6, pbh, pbw), np.uint8)}
out = {k: v.array for k, v in pins.items()}
ctx = capi.Context(0)
ctx.seq_reserve(W, H, FRAMES)
ctx.seq_upload(pin.array)
ctx.sync()


def digest():
    h = hashlib.sha1()
    h.update(out["blocks"].tobytes())
    h.update(out["motion"].tobytes())
    h.update(out["n_unique"].tobytes())
    h.update(out["planes"].tobytes())
    return h.hexdigest()[:12]


configs = [(1, 0, 0)]
for lanes in (2, 4):
    for ri in (0, 8, 12, 16, 24, 32):
        for rk in (0, 4, 8, 16):
            configs.append((lanes, ri, rk))
if len(sys.argv) > 4:
    configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[4:]]
ref = None
for lanes, ri, rk in configs:
    ctx.set_schedule(lanes, ri, rk)
    for _ in range(2):
        ctx.seq_encode(0, FRAMES, SA, THR, GOP)
    ctx.sync()
    ms = []
    for _ in range(3):
        ctx.seq_encode(0, FRAMES, SA, THR, GOP)
        ms.append(ctx.last_encode_ms("total"))
    ctx.encode_sequence(pin.array, SA, THR, GOP, out=out)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.encode_sequence(pin.array, SA, THR, GOP, out=out)
    e2e = (time.perf_counter() - t0) / 3 * 1e3
    d = digest()
    if ref is None:
        ref = d
    print(f"lanes {lanes} rows_intra {ri:3d} rows_inter {rk:3d}: resident {np.mean(ms):7.2f} ms  e2e {e2e:7.2f} ms  "
          f"stages fit {ctx.last_encode_ms('fit'):.2f} inter {ctx.last_encode_ms('inter'):.2f} intra {ctx.last_encode_ms('intra'):.2f}  "
          f"{'OK' if d == ref else 'MISMATCH ' + d}", flush=True)
