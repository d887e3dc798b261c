"""compute-sanitizer target (memcheck / synccheck / initcheck): every kernel of the library on small inputs,
results checked against the oracle -- the intra wavefront with two and three CTAs per row, inter frames with
K3s and with the row wavefront taking the leftovers, word-diverse content (chunked path), the pixel-granular
search, the decoder with bulk-copy stores, the stream path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_sequence  # noqa: E402
from oracle import port  # noqa: E402


def check(ctx, frames, sa, thr, gop, tag):
    out = ctx.encode_sequence(frames, sa, thr, gop)
    prev = None
    for i in range(frames.shape[0]):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        prev = blocks
        assert np.array_equal(out["blocks"][i], blocks) and np.array_equal(out["motion"][i], motion), (tag, i)
    print("ok", tag)
    return out


ctx = capi.Context(0)
check(ctx, make_sequence(320, 128, 4, seed=3), 16, 50, 2, "320x128 sa16")
check(ctx, make_sequence(512, 64, 3, seed=4), 4, 20, 3, "512x64 sa4 (rows of 4 groups)")
check(ctx, make_sequence(320, 128, 3, seed=8), 16, 300, 3, "320x128 sa16 thr300 (K2: int16 table)")
check(ctx, make_sequence(200, 72, 3, seed=9), 16, 40000, 3, "200x72 sa16 thr40000 (K2: round-1 tiling, int32 table; ragged tiles)")
check(ctx, make_sequence(320, 128, 2, seed=10), 24, 50, 2, "320x128 sa24 (K2: window too large for the wide kernel)")
rng = np.random.default_rng(5)
noise = rng.integers(0, 256, size=(3, 96, 256, 3), dtype=np.uint8)
check(ctx, noise, 16, 0, 3, "noise thr0 (chunked + direct paths)")
os.environ["MPTC_SPARSE_MAX_PCT"] = "0"
ctx2 = capi.Context(0)
check(ctx2, make_sequence(256, 96, 3, seed=6), 8, 0, 3, "row wavefront takes inter leftovers")
fr = make_sequence(128, 96, 2, seed=11)
enc = ctx.encode_sequence(fr[:1], 4, 50, 1)
got = ctx.inter_pixel_search(fr[1], 6, enc["blocks"][0].copy())
want = port.inter_pixel_search(fr[1], 6, port.dxt1_fit(fr[1]), enc["blocks"][0])
assert all(np.array_equal(got[k], want[k]) for k in want)
print("ok inter pixel search")
frames = make_sequence(256, 256, 4, seed=77)
stream, _ = capi.encode_stream(ctx, frames, 4, 50, 2, threads=2)
blocks, rgb, _ = capi.decode_stream(ctx, stream, threads=2, rgb=True)
assert np.array_equal(port.decode_stream(stream), blocks)
assert np.array_equal(rgb[0], port.decode_rgb(blocks[0], 256, 256))
print("ok stream round trip")
