"""Two batches in flight: steps alternate between two contexts on the same GPU (each with its own device buffers
and streams), so one batch's kernel tails, short kernels and latency-bound intra frames overlap the other's inter search.
Resident inputs; device time of K steps = wall time between the first enqueue and the last sync.
usage: python profiles/micro/two_contexts.py [n_contexts ...]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 60, 16, 50, 15
frames = np.stack([make_frame(W, H, f) for f in range(N)])
K = 12
for n_ctx in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    ctxs = [capi.Context(0) for _ in range(n_ctx)]
    for c in ctxs:
        c.seq_reserve(W, H, N)
        c.seq_upload(frames)
        c.seq_encode(0, N, SA, THR, GOP)
        c.sync()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(K):
            ctxs[i % n_ctx].seq_encode(0, N, SA, THR, GOP)
        for c in ctxs:
            c.sync()
        best = min(best, (time.perf_counter() - t0) * 1e3 / K)
    outs = [c.seq_download(0, N, want=("blocks", "motion")) for c in ctxs]
    hs = {hashlib.sha256(o["blocks"].tobytes() + o["motion"].tobytes()).hexdigest()[:16] for o in outs}
    print(f"{n_ctx} context(s): {best:.3f} ms per step of {N} frames ({W * H * N / best / 1e3:.0f} Mpixel/s), results {hs}")
    del ctxs
