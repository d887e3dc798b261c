"""GPU: the CUDA path (through the C-ABI) against the committed golden fixtures produced by the
unmodified reference, and -- at BASELINE.json's full sizes -- through size-independent
properties: decoder-side reconstruction round trip, the per-block inductive check against the
oracle on sampled blocks, determinism, PSNR of the decoded blocks."""
import numpy as np
import pytest

from golden_util import check_sequence_against_golden, load, sequence_cases, sha
from mptc_b200.synth import make_frame, make_sequence
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sequence_cases())
def test_cuda_matches_reference_fixture(ctx, name):
    g = load(name)
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    assert sha(frames) == str(g["frames_sha"])
    ctx.seq_reserve(w, h, n)
    ctx.seq_upload(frames)
    ctx.seq_encode(0, n, sa, thr, gop)
    out = ctx.seq_download(0, n)
    res = []
    for i in range(n):
        nu = int(out["n_unique"][i])
        r = {"initial": out["initial"][i], "blocks": out["blocks"][i], "motion": out["motion"][i],
             "unique": out["unique"][i, :nu]}
        if w % 256 == 0 and h % 256 == 0:
            r["planes"] = out["planes"][i].reshape(6, -1)
        assert abs(port.psnr(frames[i], out["blocks"][i]) - float(g[f"psnr_physical_{i}"])) < 0.01  # dB
        res.append(r)
    check_sequence_against_golden(g, res)


def _full_size_properties(ctx, w, h, n, sa, thr, gop, samples_per_frame, check_frames):
    frames = np.stack([make_frame(w, h, f) for f in range(n)])
    bw, bh = w // 4, h // 4
    out = ctx.encode_sequence(frames, sa, thr, gop)
    again = ctx.encode_sequence(frames, sa, thr, gop)
    for k in ("blocks", "motion", "n_unique", "planes"):
        assert np.array_equal(out[k], again[k]), f"{k} not deterministic"
    rng = np.random.default_rng(0)
    prev_words = None
    for i in range(n):
        intra = i % gop == 0
        blocks, motion = out["blocks"][i], out["motion"][i]
        nu = int(out["n_unique"][i])
        words = (blocks >> np.uint64(32)).astype(np.uint32)
        # decoder round trip: motion + unique list + previous words rebuild every index word
        rec, used = port.reconstruct_words(motion, out["unique"][i, :nu], None if intra else prev_words, bw, bh, sa)
        assert used == nu
        assert np.array_equal(rec, words), f"frame {i}: reconstruction differs"
        m2 = motion.reshape(-1, 2)
        assert ((m2[:, 0] == 255) & (m2[:, 1] == 255)).sum() == nu
        if intra:
            assert not (((m2[:, 0] & 0x80) != 0) & (m2[:, 0] != 255)).any(), "inter vector in an intra frame"
        # planes equal the oracle's planes of the same blocks
        if i in check_frames:
            assert np.array_equal(out["planes"][i], port.endpoint_planes(blocks, bw, bh))
            init = port.dxt1_fit(frames[i])
            which = np.concatenate([rng.integers(0, bw * bh, samples_per_frame),
                                    np.arange(0, min(bw * bh, 2 * bw)), np.arange(bw * bh - bw, bw * bh)])
            bad = port.check_blocks(frames[i], intra, sa, thr, init, blocks, None if intra else out["blocks"][i - 1],
                                    motion, which)
            assert bad == 0, f"frame {i}: {bad} of {which.size} sampled blocks differ from the oracle's decision"
            assert port.psnr(frames[i], blocks) > 30.0
        prev_words = words


def test_1080p_properties(ctx):
    """BASELINE configs[1] geometry (1920x1080, sa=16, thr=50): two GOP starts + inter frames."""
    _full_size_properties(ctx, 1920, 1080, 4, 16, 50, 2, samples_per_frame=1500, check_frames=(0, 1, 3))


def test_4k_properties(ctx):
    """BASELINE configs[2] geometry (3840x2160, sa=16)."""
    _full_size_properties(ctx, 3840, 2160, 2, 16, 50, 2, samples_per_frame=1200, check_frames=(0, 1))


@pytest.mark.parametrize("sa,thr", [(2, 0), (4, 10), (8, 200), (32, 50)])
def test_window_threshold_sweep(ctx, sa, thr):
    """BASELINE configs[3]: search window / threshold sweep, full oracle comparison at a size the
    oracle finishes in seconds."""
    w, h, n, gop = 192, 128, 3, 3
    frames = make_sequence(w, h, n, seed=sa * 100 + thr)
    out = ctx.encode_sequence(frames, sa, thr, gop)
    prev = None
    for i in range(n):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i == 0, sa, thr, init, prev)
        assert np.array_equal(out["blocks"][i], blocks), f"sa={sa} thr={thr} frame {i}"
        assert np.array_equal(out["motion"][i], motion)
        assert np.array_equal(out["unique"][i, : out["n_unique"][i]], unique)
        prev = blocks


def test_ragged_last_gop_and_gop1(ctx):
    """Frame count not a multiple of the GOP; GOP of 1 (all intra)."""
    frames = make_sequence(128, 64, 5, seed=31)
    for gop in (1, 2, 3):
        out = ctx.encode_sequence(frames, 4, 50, gop)
        prev = None
        for i in range(5):
            init = port.dxt1_fit(frames[i])
            blocks, motion, unique = port.reencode(frames[i], i % gop == 0, 4, 50, init, prev)
            assert np.array_equal(out["blocks"][i], blocks), f"gop={gop} frame {i}"
            assert np.array_equal(out["motion"][i], motion)
            prev = blocks


def test_bad_arguments_are_errors(ctx):
    from mptc_b200.capi import MptcError
    with pytest.raises(MptcError):
        ctx.encode_sequence(np.zeros((1, 30, 32, 3), np.uint8), 4, 50, 1)   # height not a multiple of 4
    with pytest.raises(MptcError):
        ctx.encode_sequence(np.zeros((1, 32, 32, 3), np.uint8), 64, 50, 1)  # search_area > 63
