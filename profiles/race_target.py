"""compute-sanitizer --tool racecheck target: small frames, intra + inter + leftovers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H = 512, 192
frames = np.stack([make_frame(1920, 1080, f)[256:256 + H, 640:640 + W] for f in (0, 1, 15, 16)])
frames = np.ascontiguousarray(frames)
ctx = capi.Context(0)
out = ctx.encode_sequence(frames, 16, 10, 2)
print(out["n_unique"])
