"""GPU: the decoder side through the C-ABI -- index words by pointer jumping,
inverse endpoint planes, block assembly, DXT1 -> RGB, and the whole-stream decode -- bit-exact
against the oracle and against what the reference's own decoder produced (fixtures)."""
import numpy as np
import pytest

from golden_util import load, sha
from mptc_b200 import capi
from mptc_b200.synth import make_sequence
from oracle import port

pytestmark = pytest.mark.gpu


def test_reference_stream_decodes_to_reference_decoder_output(ctx):
    """The stream CompressMultiUnique wrote -> the blocks and pictures the reference decoder made."""
    g = load("stream_256x256_sa4_gop2")
    d = load("decode_256x256_sa4_gop2")
    for threads in (1, 6):
        blocks, rgb, st = capi.decode_stream(ctx, g["stream"].tobytes(), threads=threads, rgb=True)
        assert np.array_equal(blocks, d["blocks"])
        for i in range(blocks.shape[0]):
            assert sha(rgb[i]) == str(d["rgb_sha"][i]), f"frame {i}"
        assert np.array_equal(rgb[-1, :16], d["rgb_last_rows"])
        assert st.symbols > 0 and st.header.n_frames == blocks.shape[0]


@pytest.mark.parametrize("w,h,n,sa,thr,gop,seed", [
    (192, 128, 7, 4, 30, 3, 21),      # planes padded to 64 (extension), ragged last GOP
    (256, 256, 4, 16, 50, 4, 1234),   # BASELINE configs[0] geometry
    (64, 64, 3, 1, 200, 1, 3),        # intra only, smallest search area
    (320, 192, 6, 8, 0, 2, 9),        # thr 0
    (132, 68, 4, 3, 10, 2, 2),        # odd block counts (33 x 17)
])
def test_encode_then_decode_round_trip(ctx, w, h, n, sa, thr, gop, seed):
    """GPU encoder results -> GPU decoder: the encoder's final blocks come back exactly; each stage
    equals the oracle's restatement of the reference decoder."""
    frames = make_sequence(w, h, n, seed=seed)
    enc = ctx.encode_sequence(frames, sa, thr, gop)
    blocks, rgb = ctx.decode_sequence(enc["motion"], enc["unique"], enc["n_unique"], enc["planes"], w, h, sa, gop, rgb=True)
    assert np.array_equal(blocks, enc["blocks"])
    bw, bh = w // 4, h // 4
    prev = None
    for i in range(n):
        nu = int(enc["n_unique"][i])
        words, used = port.reconstruct_words(enc["motion"][i], enc["unique"][i, :nu], prev if i % gop else None, bw, bh, sa)
        assert used == nu
        ep1, ep2 = port.inverse_planes(enc["planes"][i], bw, bh)
        want = ep1.astype(np.uint64) | (ep2.astype(np.uint64) << np.uint64(16)) | (words.astype(np.uint64) << np.uint64(32))
        assert np.array_equal(blocks[i], want), f"frame {i}"
        assert np.array_equal(rgb[i], port.decode_rgb(want, w, h)), f"rgb {i}"
        prev = words
    # packed unique layout (the stream's group palettes) gives the same result
    packed = np.concatenate([enc["unique"][i, : int(enc["n_unique"][i])] for i in range(n)] + [np.zeros(1, np.uint32)])
    again = ctx.decode_sequence(enc["motion"], packed, enc["n_unique"], enc["planes"], w, h, sa, gop, packed=True)
    assert np.array_equal(again, enc["blocks"])


def test_long_copy_chains(ctx):
    """Flat content: every block copies its left neighbour, frame after frame -- chains thousands of
    links long, collapsed by pointer jumping."""
    w, h, n, sa, gop = 512, 256, 6, 2, 6
    frames = np.full((n, h, w, 3), 77, dtype=np.uint8)
    frames[:, :4, :4] = 200      # one distinct block so there are two unique words
    enc = ctx.encode_sequence(frames, sa, 50, gop)
    assert int(enc["n_unique"].sum()) <= 8
    blocks = ctx.decode_sequence(enc["motion"], enc["unique"], enc["n_unique"], enc["planes"], w, h, sa, gop)
    assert np.array_equal(blocks, enc["blocks"])


def test_arbitrary_symbols_match_oracle(ctx):
    """Random (valid) motion fields and random wavelet symbols, not produced by any encoder: the
    kernels still equal the oracle bit for bit (int8 wrap-arounds of the inverse transform included)."""
    rng = np.random.default_rng(12)
    w, h, n, sa, gop = 160, 96, 5, 5, 5
    bw, bh = w // 4, h // 4
    nb = bw * bh
    pbw, pbh = 64, 64
    motion = np.zeros((n, nb, 2), dtype=np.uint8)
    n_unique = np.zeros(n, dtype=np.uint32)
    unique = rng.integers(0, 2**32, (n, nb), dtype=np.uint64).astype(np.uint32)
    for f in range(n):
        for b in range(nb):
            bx, by = b % bw, b // bw
            kind = rng.integers(0, 3)
            if kind == 1 and f % gop:
                rx, ry = rng.integers(max(bx - sa, 0), min(bx + sa, bw)), rng.integers(max(by - sa, 0), min(by + sa, bh))
                motion[f, b] = ((rx - bx + sa) | 0x80, (ry - by + sa) | 0x80)
            elif kind == 2 and b > 0:
                ry = rng.integers(max(by - 2 * sa + 1, 0), by + 1)
                lo, hi = max(bx - sa, 0), (bx if ry == by else min(bx + sa, bw))
                if hi > lo:
                    rx = rng.integers(lo, hi)
                    motion[f, b] = (rx - bx + sa, ry - by + 2 * sa - 1)
                    continue
                motion[f, b] = (255, 255)
                n_unique[f] += 1
            else:
                motion[f, b] = (255, 255)
                n_unique[f] += 1
    planes = rng.integers(96, 160, (n, 6, pbh, pbw)).astype(np.uint8)
    blocks = ctx.decode_sequence(motion.reshape(n, -1), unique, n_unique, planes, w, h, sa, gop)
    prev = None
    for f in range(n):
        words, used = port.reconstruct_words(motion[f].reshape(-1), unique[f, : n_unique[f]], prev if f % gop else None, bw, bh, sa)
        assert used == n_unique[f]
        ep1, ep2 = port.inverse_planes(planes[f], bw, bh)
        want = ep1.astype(np.uint64) | (ep2.astype(np.uint64) << np.uint64(16)) | (words.astype(np.uint64) << np.uint64(32))
        assert np.array_equal(blocks[f], want), f"frame {f}"
        prev = words


def test_corrupt_motion_is_reported(ctx):
    w, h, sa = 64, 64, 4
    nb = 256
    motion = np.full((1, 2 * nb), 255, dtype=np.uint8)
    motion[0, 0:2] = (sa, 2 * sa - 1)          # block 0 copying itself: never emitted
    unique = np.arange(nb, dtype=np.uint32).reshape(1, nb)
    planes = np.full((1, 6, 64, 64), 128, dtype=np.uint8)
    with pytest.raises(capi.MptcError, match="corrupt"):
        ctx.decode_sequence(motion, unique, np.array([nb - 1], np.uint32), planes, w, h, sa, 1)
    motion[0, 0:2] = (0x80 | sa, 0x80 | sa)    # inter vector in an intra frame
    with pytest.raises(capi.MptcError, match="corrupt"):
        ctx.decode_sequence(motion, unique, np.array([nb - 1], np.uint32), planes, w, h, sa, 1)
    # and the context still works afterwards
    motion[0, 0:2] = (255, 255)
    blocks = ctx.decode_sequence(motion, unique, np.array([nb], np.uint32), planes, w, h, sa, 1)
    assert np.array_equal((blocks[0] >> np.uint64(32)).astype(np.uint32), unique[0])


def test_full_size_device_resident_round_trip(ctx):
    """1080p, sa 16, one GOP of 3 frames: decode straight from what the encoder left on the device;
    the decoded blocks equal the encoder's final blocks and the picture's PSNR equals the PSNR of
    the encoder's blocks (0.01 dB bar of the north star; here it is exact)."""
    w, h, n, sa, thr, gop = 1920, 1080, 3, 16, 50, 3
    frames = make_sequence(w, h, n, seed=1234)
    ctx.seq_reserve(w, h, n)
    ctx.seq_upload(frames)
    ctx.seq_encode(0, n, sa, thr, gop)
    ctx.seq_decode(0, n, sa, gop, rgb=True)
    enc = ctx.seq_download(0, n, want=("blocks",))
    blocks, rgb = ctx.seq_decode_download(0, n, rgb=True)
    assert np.array_equal(blocks, enc["blocks"])
    assert ctx.last_decode_ms("total") > 0
    for i in range(n):
        mse = np.mean((frames[i].astype(np.float64) - rgb[i]) ** 2)
        psnr = 10 * np.log10(255.0 ** 2 / mse)
        assert abs(psnr - port.psnr(frames[i], blocks[i])) < 1e-6
    assert np.array_equal(rgb[1], port.decode_rgb(blocks[1], w, h))


def test_stream_round_trip_through_both_host_coders(ctx):
    """encode_stream -> decode_stream at a size with padded planes (extension) and 3 GOPs."""
    w, h, n, sa, thr, gop = 384, 192, 6, 8, 50, 2
    frames = make_sequence(w, h, n, seed=8)
    stream, _ = capi.encode_stream(ctx, frames, sa, thr, gop, threads=4)
    enc = ctx.encode_sequence(frames, sa, thr, gop)
    blocks, rgb, st = capi.decode_stream(ctx, stream, threads=4, rgb=True)
    assert np.array_equal(blocks, enc["blocks"])
    assert np.array_equal(rgb[n - 1], port.decode_rgb(blocks[n - 1], w, h))
    with pytest.raises(capi.MptcError):
        capi.decode_stream(ctx, stream[: len(stream) - 100])
