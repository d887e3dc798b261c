"""GPU: K2p, the pixel-granular inter search (DXTImage::InterPixelSearch, dxt_image.cpp:776-832 with the
pattern of DXTImage::SetPattern, dxt_image.h:135-164 -- SURVEY.md 8(f)-4), through the C ABI
(mptc_gpu_inter_pixel_search), bit-exact against
  * the committed fixture produced by the reference's own CompressedBlock methods (the function's loop
    with the undefined behaviour of Get4X4InterpolationBlock removed, tests/golden/gen_golden.py), and
  * the oracle's restatement on other geometries (frame borders, search areas 1 .. 40, static and moving
    content, the previous frame's blocks taken from a real encode)."""
import numpy as np
import pytest

from golden_util import load
from mptc_b200.synth import make_sequence
from oracle import port

pytestmark = pytest.mark.gpu

KEYS = ("min_err", "motion", "index", "reassigned")


def test_matches_the_reference_fixture(ctx):
    g = load("inter_pixel_search")
    for name in ("a", "b"):
        w, h, seed, sa = [int(x) for x in g[f"{name}_params"]]
        fr = make_sequence(w, h, 2, seed=seed)
        got = ctx.inter_pixel_search(fr[1], sa, g[f"{name}_prev"])        # cur_blocks = the stb fit, on the device
        for k in KEYS:
            assert np.array_equal(got[k], g[f"{name}_{k}"]), (name, k)
        again = ctx.inter_pixel_search(fr[1], sa, g[f"{name}_prev"], cur_blocks=g[f"{name}_cur"])
        for k in KEYS:
            assert np.array_equal(again[k], got[k]), (name, k)


@pytest.mark.parametrize("w,h,sa,seed", [(64, 48, 1, 1), (64, 48, 2, 2), (200, 120, 7, 3), (128, 128, 16, 4), (96, 96, 40, 5),
                                          (256, 64, 63, 6)])
def test_matches_the_oracle(ctx, w, h, sa, seed):
    fr = make_sequence(w, h, 2, seed=seed)
    enc = ctx.encode_sequence(fr[:1], min(sa, 16), 50, 1)                 # a real previous frame: re-assigned words
    prev = enc["blocks"][0].copy()
    want = port.inter_pixel_search(fr[1], sa, port.dxt1_fit(fr[1]), prev)
    got = ctx.inter_pixel_search(fr[1], sa, prev)
    for k in KEYS:
        assert np.array_equal(got[k], want[k]), k
    assert (got["min_err"] == 0).any()


def test_noise_and_flat_content(ctx):
    rng = np.random.default_rng(12)
    w, h, sa = 96, 64, 6
    fr = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    fr[:, 16:40, 24:72] = 77                                               # flat: all-equal index words (den == 0)
    prev = port.dxt1_fit(fr[0])
    want = port.inter_pixel_search(fr[1], sa, port.dxt1_fit(fr[1]), prev)
    got = ctx.inter_pixel_search(fr[1], sa, prev)
    for k in KEYS:
        assert np.array_equal(got[k], want[k]), k
