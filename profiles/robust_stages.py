"""Per-stage device time (one lane: kernels serialised) of the word-diverse corners: clean content at
err_threshold 0, noisy content (+-24) at 50, search_area 32.  One GOP of 15 1080p frames each."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, GOP = 1920, 1080, 15, 15
clean = np.stack([make_frame(W, H, f) for f in range(N)])
rng = np.random.default_rng(1000)
noisy = np.clip(clean.astype(np.int16) + rng.integers(-24, 25, size=clean.shape, dtype=np.int16), 0, 255).astype(np.uint8)
ctx = capi.Context(0)
ctx.seq_reserve(W, H, N)
cases = [("clean sa16 thr50", clean, 16, 50), ("clean sa16 thr0", clean, 16, 0), ("noisy sa16 thr50", noisy, 16, 50), ("clean sa32 thr50", clean, 32, 50)]
if len(sys.argv) > 1:
    cases = [c for c in cases if any(a in c[0] for a in sys.argv[1:])]
for name, frames, sa, thr in cases:
    ctx.seq_upload(frames)
    for lanes in (1, 0):
        ctx.set_schedule(lanes, 0, 0)
        for _ in range(2):
            ctx.seq_encode(0, N, sa, thr, GOP)
        ctx.sync()
        st = {k: round(ctx.last_encode_ms(k), 2) for k in capi.STAGES}
        wk = ctx.last_work_count()
        print(f"{name:18s} lanes {lanes}: {N * W * H / st['total'] / 1e3:7.0f} Mpixel/s  {st}  words/tile {wk['inter_evals'] / max(1, 32 * wk['inter_tiles']):.0f} "
              f"intra groups {wk['intra_groups']} intra evals {wk['intra_evals'] / 1e6:.0f}M")
    out = ctx.seq_download(0, N, want=("motion", "unique"))
    m = out["motion"].reshape(N, -1, 2)
    uniq = (m[:, :, 0] == 255) & (m[:, :, 1] == 255)
    inter = ((m[:, :, 0] & 0x80) != 0) & ~uniq
    print(f"{'':18s} per frame: inter-found {inter.sum(1)[1:].mean() / m.shape[1] * 100:.1f}%  unique {uniq.sum(1)[1:].mean() / m.shape[1] * 100:.1f}%  "
          f"intra-found (inter frames) {100 - (inter.sum(1)[1:].mean() + uniq.sum(1)[1:].mean()) / m.shape[1] * 100:.1f}%")
