"""A/B of a kernel change on the headline workload: device ms per 60-frame 1080p step (best of 5),
one-lane stage times, and a hash of the results (must not change).
usage: python profiles/micro/k2_ab.py"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 60, int(os.environ.get("SA", "16")), int(os.environ.get("THR", "50")), 15
pin = capi.PinnedArray((N, H, W, 3), np.uint8)
for f in range(N):
    pin.array[f] = make_frame(W, H, f)
ctx = capi.Context(0)
ctx.seq_reserve(W, H, N)
ctx.seq_upload(pin.array)
best = 1e9
for _ in range(8):
    ctx.seq_encode(0, N, SA, THR, GOP)
    best = min(best, ctx.last_encode_ms("total"))
out = ctx.seq_download(0, N, want=("blocks", "motion"))
h = hashlib.sha256(out["blocks"].tobytes() + out["motion"].tobytes()).hexdigest()[:16]
ctx.set_schedule(1, 0, 0)
st = None
for _ in range(4):
    ctx.seq_encode(0, N, SA, THR, GOP)
    t = {k: round(ctx.last_encode_ms(k), 3) for k in ("total", "fit", "inter", "intra")}
    st = t if st is None or t["total"] < st["total"] else st
print(f"K2={os.environ.get('MPTC_K2', 'default')} sa {SA} thr {THR}: step {best:.3f} ms  one-lane {st}  results {h}  work {ctx.last_work_count()}")
