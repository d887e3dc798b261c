"""compute-sanitizer target for the paths added late in round 1: word-diverse content (K3 chunked
path, K2 multi-chunk), the two-pass compaction and the decoder kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402

rng = np.random.default_rng(3)
W, H, N, SA, THR, GOP = 192, 96, 3, 8, 0, 3
frames = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
frames[:, :, W // 2:] = (frames[:, :, W // 2:] // 64) * 64      # a calmer half: groups on both paths
ctx = capi.Context(0)
out = ctx.encode_sequence(frames, SA, THR, GOP)
blocks, rgb = ctx.decode_sequence(out["motion"], out["unique"], out["n_unique"], out["planes"], W, H, SA, GOP, rgb=True)
assert np.array_equal(blocks, out["blocks"])
print("ok", out["n_unique"], int(rgb.sum()) % 1000)
