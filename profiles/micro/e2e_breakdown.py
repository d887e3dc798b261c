"""End-to-end step (pinned host frames -> host results): device time between the call's first and
last event against the wall time of the call, and against the resident step."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 60, 16, 50, 15
nb = (W // 4) * (H // 4)
pin = capi.PinnedArray((N, H, W, 3), np.uint8)
for f in range(N):
    pin.array[f] = make_frame(W, H, f)
pbw, pbh = (W // 4 + 63) // 64 * 64, (H // 4 + 63) // 64 * 64
pins = {"blocks": capi.PinnedArray((N, nb), np.uint64), "motion": capi.PinnedArray((N, 2 * nb), np.uint8),
        "unique": capi.PinnedArray((N, nb), np.uint32), "n_unique": capi.PinnedArray((N,), np.uint32),
        "planes": capi.PinnedArray((N, 6, pbh, pbw), np.uint8)}
out = {k: v.array for k, v in pins.items()}
ctx = capi.Context(0)
for it in range(6):
    t0 = time.perf_counter()
    ctx.encode_sequence(pin.array, SA, THR, GOP, out=out)
    wall = (time.perf_counter() - t0) * 1e3
    t1 = time.perf_counter()
    ctx.encode_sequence(pin.array, SA, THR, GOP, out=out, wait=False)
    enq = (time.perf_counter() - t1) * 1e3
    ctx.wait()
    print(f"e2e wall {wall:.2f} ms, device (events) {ctx.last_encode_ms('total'):.2f} ms, host enqueue alone {enq:.2f} ms")
ctx.seq_upload(pin.array)
for it in range(3):
    ctx.seq_encode(0, N, SA, THR, GOP)
    print(f"resident device {ctx.last_encode_ms('total'):.2f} ms")
