"""Short workload for ncu captures: 1080p, 2 GOPs x 2 frames, sa=16 (same geometry and
parameters as bench.py's workload, fewer frames so that --set full replays stay short)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 4, 16, 50, 2
frames = np.stack([make_frame(W, H, f) for f in (0, 1, 15, 16)])
ctx = capi.Context(0)
ctx.seq_reserve(W, H, N)
ctx.seq_upload(frames)
for _ in range(2):
    ctx.seq_encode(0, N, SA, THR, GOP)
    ctx.sync()
print({k: round(ctx.last_encode_ms(k), 3) for k in capi.STAGES}, ctx.last_candidate_count())
for _ in range(2):   # decoder side: from the symbols the encode left on the device
    ctx.seq_decode(0, N, SA, GOP, rgb=True)
    ctx.sync()
print({k: round(ctx.last_decode_ms(k), 3) for k in capi.DECODE_STAGES})
