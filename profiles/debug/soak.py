"""Determinism soak: the benchmarked step repeated, every result hashed (blocks, motion, unique counts)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 60, 16, 50, 15
ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 100
frames = np.stack([make_frame(W, H, f) for f in range(N)])
ctx = capi.Context(0)
ctx.seq_reserve(W, H, N)
ctx.seq_upload(frames)
seen = {}
for it in range(ITER):
    ctx.seq_encode(0, N, SA, THR, GOP)
    out = ctx.seq_download(0, N, want=("blocks", "motion", "unique"))
    h = hashlib.sha256(out["blocks"].tobytes() + out["motion"].tobytes() + out["n_unique"].tobytes()).hexdigest()[:16]
    seen[h] = seen.get(h, 0) + 1
print(f"{ITER} encodes of {N} x {W}x{H}: distinct results {seen}")
rng = np.random.default_rng(3)
noise = rng.integers(0, 256, (6, 256, 512, 3), dtype=np.uint8)
seen = {}
for it in range(ITER):
    out = ctx.encode_sequence(noise, 16, (0, 50, 300)[it % 3], 3)
    h = hashlib.sha256(out["blocks"].tobytes() + out["motion"].tobytes()).hexdigest()[:16]
    seen[(it % 3, h)] = seen.get((it % 3, h), 0) + 1
print(f"{ITER} encodes of noise at three thresholds: distinct results per threshold {seen}")
