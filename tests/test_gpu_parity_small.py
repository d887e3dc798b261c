"""GPU parity against the oracle on small seeded inputs (bit-exact: integer/byte/index work).
Calls go through the C-ABI (mptc_b200.capi -> libmptc_b200.so)."""
import numpy as np
import pytest

from mptc_b200.synth import make_sequence
from oracle import port

pytestmark = pytest.mark.gpu


def oracle_sequence(frames, sa, thr, gop):
    prev = None
    res = []
    for i, rgb in enumerate(frames):
        init = port.dxt1_fit(rgb)
        blocks, motion, unique = port.reencode(rgb, i % gop == 0, sa, thr, init, prev)
        res.append((init, blocks, motion, unique))
        prev = blocks
    return res


@pytest.mark.parametrize("w,h", [(64, 64), (256, 256), (200, 120)])
def test_dxt1_fit_bit_exact(ctx, w, h):
    frames = make_sequence(w, h, 2, seed=7)
    for rgb in frames:
        assert np.array_equal(ctx.dxt1_fit(rgb), port.dxt1_fit(rgb))


def test_dxt1_fit_edge_content(ctx):
    rng = np.random.default_rng(3)
    h, w = 64, 128
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)           # noise
    img[:16] = 0                                                          # flat black
    img[16:32] = 255                                                      # flat white
    img[32:36, :, :] = rng.integers(0, 256, size=(1, w, 3), dtype=np.uint8)  # vertical stripes
    img[36:40, :, 1:] = 0                                                 # single channel
    img[40:44] = np.arange(w, dtype=np.uint8)[None, :, None]             # grey ramp
    assert np.array_equal(ctx.dxt1_fit(img), port.dxt1_fit(img))


@pytest.mark.parametrize("sa,thr,gop", [(2, 50, 2), (4, 10, 3), (8, 50, 4), (16, 50, 4), (4, 0, 2), (3, 200, 2)])
def test_reencode_single_frame_api(ctx, sa, thr, gop):
    frames = make_sequence(128, 96, gop, seed=11)
    ref = oracle_sequence(frames, sa, thr, gop)
    prev = None
    for i, rgb in enumerate(frames):
        got = ctx.reencode(rgb, i == 0, sa, thr, prev)
        init, blocks, motion, unique = ref[i]
        assert np.array_equal(got["initial"], init)
        assert np.array_equal(got["motion"], motion), f"frame {i}"
        assert np.array_equal(got["blocks"], blocks), f"frame {i}"
        assert np.array_equal(got["unique"], unique), f"frame {i}"
        prev = got["blocks"]


@pytest.mark.parametrize("w,h,n,sa,thr,gop", [(256, 256, 8, 8, 50, 4), (256, 256, 8, 16, 50, 4), (320, 192, 6, 4, 20, 3)])
def test_sequence_bit_exact(ctx, w, h, n, sa, thr, gop):
    frames = make_sequence(w, h, n)
    ref = oracle_sequence(frames, sa, thr, gop)
    out = ctx.encode_sequence(frames, sa, thr, gop)
    for i in range(n):
        init, blocks, motion, unique = ref[i]
        assert np.array_equal(out["motion"][i], motion), f"frame {i}"
        assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
        nu = int(out["n_unique"][i])
        assert nu == unique.size
        assert np.array_equal(out["unique"][i, :nu], unique)
        planes = port.endpoint_planes(blocks, w // 4, h // 4)
        assert np.array_equal(out["planes"][i], planes), f"planes frame {i}"


def test_flat_content_den_zero_path(ctx):
    """Flat frames give all-equal index words (den == 0 in RecalculateEndpoints)."""
    frames = np.empty((2, 64, 64, 3), dtype=np.uint8)
    frames[0] = 77
    frames[1] = 77
    frames[1, 8:24, 8:40] = (10, 200, 30)
    frames[0, 40:44, 0:64, 0] = np.arange(64, dtype=np.uint8) * 3
    ref = oracle_sequence(frames, 4, 50, 2)
    out = ctx.encode_sequence(frames, 4, 50, 2)
    for i in range(2):
        assert np.array_equal(out["blocks"][i], ref[i][1])
        assert np.array_equal(out["motion"][i], ref[i][2])


def test_endpoint_planes_api(ctx):
    rng = np.random.default_rng(5)
    for bw, bh in [(64, 64), (128, 64), (100, 70)]:
        blocks = rng.integers(0, 2**63, size=bw * bh, dtype=np.uint64)
        assert np.array_equal(ctx.endpoint_planes(blocks, bw, bh), port.endpoint_planes(blocks, bw, bh))


def _diverse_frames(w, h, n, seed, kind):
    """Content whose blocks keep ~distinct index words: the word tables of the tiled kernels
    overflow and the chunked paths run."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    base = make_sequence(w, h, n, seed=seed).astype(np.int16)
    if kind == "noisy":        # the synthetic sequence + strong noise (camera-like)
        base += rng.integers(-24, 25, size=base.shape, dtype=np.int16)
    else:                      # "mixed": noise on the left half, clean content on the right
        base[:, :, : w // 2] += rng.integers(-40, 41, size=(n, h, w // 2, 3), dtype=np.int16)
    return np.clip(base, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("w,h,n,sa,thr,gop,kind", [
    (256, 128, 3, 16, 0, 3, "noise"),      # every window far beyond the 256-word table, intra + inter leftovers
    (256, 128, 2, 16, 50, 2, "noisy"),
    (384, 96, 3, 8, 0, 3, "mixed"),        # groups on the fast path next to overflowing ones
    (192, 192, 2, 16, 10, 1, "noisy"),     # intra only (two CTAs per row)
    (132, 100, 3, 12, 0, 2, "noise"),      # ragged group at the row end, odd block counts
    (192, 128, 2, 24, 0, 2, "noise"),      # window wider than 32: the general scan path, chunked
    (256, 64, 3, 3, 0, 3, "noise"),        # tiny window, nearly every block left over
    (160, 160, 2, 40, 5, 2, "mixed"),      # search area larger than the tiled kernels' tables on noise
])
def test_word_diverse_content_bit_exact(ctx, w, h, n, sa, thr, gop, kind):
    frames = _diverse_frames(w, h, n, 17, kind)
    ref = oracle_sequence(frames, sa, thr, gop)
    out = ctx.encode_sequence(frames, sa, thr, gop)
    for i in range(n):
        init, blocks, motion, unique = ref[i]
        assert np.array_equal(out["motion"][i], motion), f"frame {i}"
        assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
        nu = int(out["n_unique"][i])
        assert nu == unique.size and np.array_equal(out["unique"][i, :nu], unique)


def test_word_diverse_content_with_row_wavefront_for_leftovers(ctx):
    """The same with K3s switched off, so that the row wavefront (and its chunked path) takes the
    inter frames' leftovers as well."""
    import os
    from mptc_b200 import capi
    os.environ["MPTC_SPARSE_MAX_PCT"] = "0"
    try:
        c2 = capi.Context(0)
    finally:
        del os.environ["MPTC_SPARSE_MAX_PCT"]
    try:
        for kind, thr in (("noise", 0), ("mixed", 0), ("noisy", 50)):
            w, h, n, sa, gop = 256, 128, 3, 16, 3
            frames = _diverse_frames(w, h, n, 23, kind)
            ref = oracle_sequence(frames, sa, thr, gop)
            out = c2.encode_sequence(frames, sa, thr, gop)
            for i in range(n):
                assert np.array_equal(out["motion"][i], ref[i][2]), f"{kind} frame {i}"
                assert np.array_equal(out["blocks"][i], ref[i][1]), f"{kind} frame {i}"
    finally:
        c2.close()


@pytest.mark.parametrize("w,h,n,sa,thr,gop", [
    (4, 4, 3, 1, 50, 2),          # one block per frame
    (8, 4, 2, 63, 50, 2),         # the largest search area the uint8 motion bytes allow, tiny frame
    (64, 32, 2, 63, 20, 2),       # search area far larger than the frame
    (128, 64, 1, 40, 50, 1),      # single frame, window > 32 (multi-step scan paths)
    (96, 64, 4, 5, -1, 2),        # negative threshold: only strictly better candidates are found
    (96, 64, 3, 5, 1 << 30, 3),   # huge threshold: everything with an accepted candidate is found
    (32, 128, 5, 7, 50, 255),     # gop longer than the sequence; tall narrow frame
])
def test_edge_geometries_and_thresholds(ctx, w, h, n, sa, thr, gop):
    frames = make_sequence(w, h, n, seed=29)
    ref = oracle_sequence(frames, sa, thr, gop)
    out = ctx.encode_sequence(frames, sa, thr, gop)
    for i in range(n):
        init, blocks, motion, unique = ref[i]
        assert np.array_equal(out["motion"][i], motion), f"frame {i}"
        assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
        nu = int(out["n_unique"][i])
        assert nu == unique.size and np.array_equal(out["unique"][i, :nu], unique)
        assert np.array_equal(out["planes"][i], port.endpoint_planes(blocks, w // 4, h // 4)), f"planes {i}"
    # and back through the decoder
    dec = ctx.decode_sequence(out["motion"], out["unique"], out["n_unique"], out["planes"], w, h, sa, gop)
    assert np.array_equal(dec, out["blocks"])


def test_bad_arguments_are_reported(ctx):
    from mptc_b200 import capi
    frames = make_sequence(64, 64, 2, seed=1)
    with pytest.raises(capi.MptcError):
        ctx.encode_sequence(frames, 64, 50, 2)            # search_area > 63
    with pytest.raises(capi.MptcError):
        ctx.encode_sequence(frames, 0, 50, 2)             # search_area < 1
    with pytest.raises(capi.MptcError):
        ctx.encode_sequence(frames, 4, 50, 0)             # gop < 1
    with pytest.raises(capi.MptcError):
        ctx.dxt1_fit(np.zeros((6, 8, 3), dtype=np.uint8))  # height not a multiple of 4
    out = ctx.encode_sequence(frames, 4, 50, 2)           # the context still works
    assert out["blocks"].shape == (2, 256)
