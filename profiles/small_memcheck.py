"""compute-sanitizer target: four 1080p intra frames (frames 0, 15, 30, 45 of the bench sequence)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H = 1920, 1080
frames = np.stack([make_frame(W, H, f) for f in (0, 15, 30, 45)])
ctx = capi.Context(0)
out = ctx.encode_sequence(frames, 16, 50, 1)
print(out["n_unique"])
