"""Which copies stretch the end-to-end step?  Device time (events) of mptc_gpu_encode_sequence with subsets of the
host outputs, against the resident step.  usage: python profiles/micro/e2e_parts.py"""
import ctypes as C
import os
import sys

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
ctx = capi.Context(0)
L = capi.load()
p = capi.Params(SA, THR, GOP)


def run(names):
    ptr = lambda k: pins[k].array.ctypes.data if k in names else None   # noqa: E731
    best = 1e9
    for _ in range(6):
        r = L.mptc_gpu_encode_sequence(ctx._p, pin.array.ctypes.data, N, W, H, C.byref(p), ptr("blocks"), ptr("motion"),
                                       ptr("unique"), ptr("n_unique"), ptr("planes"))
        assert r == 0, r
        best = min(best, ctx.last_encode_ms("total"))
    return best


for names in (("blocks", "motion", "unique", "n_unique", "planes"), ("n_unique",), ("blocks", "motion", "n_unique"),
              ("planes", "n_unique"), ("unique", "n_unique")):
    print(f"H2D + kernels + D2H of {', '.join(names):48s} {run(names):7.2f} ms")
ctx.seq_reserve(W, H, N)
ctx.seq_upload(pin.array)
best = 1e9
for _ in range(6):
    ctx.seq_encode(0, N, SA, THR, GOP)
    best = min(best, ctx.last_encode_ms("total"))
print(f"resident (no copies)                                                    {best:7.2f} ms")
# per-stage sums (over lanes and launches, overlapped) of the resident step and of the end-to-end step
stages = ("total", "fit", "inter", "intra", "compact", "planes")
ctx.seq_encode(0, N, SA, THR, GOP)
print("resident   ", {k: round(ctx.last_encode_ms(k), 2) for k in stages})
out = {k: v.array for k, v in pins.items()}
ctx.encode_sequence(pin.array, SA, THR, GOP, out=out)
print("end to end ", {k: round(ctx.last_encode_ms(k), 2) for k in stages})
