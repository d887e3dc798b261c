"""CPU: the oracle (oracle/mptc_oracle.c) against the golden fixtures generated from the
unmodified reference (tests/golden/gen_golden.py).  Bit-exact."""
import numpy as np
import pytest

from golden_util import check_sequence_against_golden, load, sequence_cases, sha
from mptc_b200.synth import make_sequence
from oracle import port


@pytest.mark.parametrize("name", sequence_cases())
def test_port_matches_reference_fixture(name):
    g = load(name)
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    assert sha(frames) == str(g["frames_sha"]), "synthetic generator drifted"
    prev = None
    res = []
    for i in range(n):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        r = {"initial": init, "blocks": blocks, "motion": motion, "unique": unique}
        if w % 256 == 0 and h % 256 == 0:
            r["planes"] = port.endpoint_planes(blocks, w // 4, h // 4).reshape(6, -1)
        assert abs(port.psnr(frames[i], blocks) - float(g[f"psnr_physical_{i}"])) < 1e-9
        res.append(r)
        prev = blocks
    check_sequence_against_golden(g, res)


def test_arith_coder_known_answers():
    g = load("arith")
    names = [k[4:] for k in g.files if k.startswith("sym_")]
    assert len(names) >= 5
    for k in names:
        assert port.arith_encode(g["sym_" + k]) == g["enc_" + k].tobytes(), k


def test_payload_sizes_from_planes():
    """Compressed size of every per-frame stream equals the reference's (codec.cpp:1115-1158)."""
    g = load("seq_256x256_sa8")
    for i in range(2):
        planes = g[f"planes_{i}"]
        sizes = g[f"sizes_{i}"]
        mine = [len(port.arith_encode(g[f"motion_{i}"])), len(port.arith_encode(planes[0])),
                len(port.arith_encode(np.concatenate([planes[1], planes[2]]))), len(port.arith_encode(planes[3])),
                len(port.arith_encode(np.concatenate([planes[4], planes[5]])))]
        assert mine == [int(s) for s in sizes]


def test_reconstruct_words_round_trip():
    """Decoder-side reconstruction (codec.cpp:441-500) of the oracle's own output."""
    w, h, n, sa, thr, gop = 128, 96, 4, 4, 20, 2
    frames = make_sequence(w, h, n, seed=21)
    prev = None
    for i in range(n):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        pw = None if prev is None else (prev >> np.uint64(32)).astype(np.uint32)
        words, used = port.reconstruct_words(motion, unique, pw, w // 4, h // 4, sa)
        assert used == unique.size
        assert np.array_equal(words, (blocks >> np.uint64(32)).astype(np.uint32))
        assert port.check_blocks(frames[i], i % gop == 0, sa, thr, init, blocks, prev, motion,
                                 np.arange(blocks.size)) == 0
        prev = blocks


def test_check_blocks_detects_corruption():
    frames = make_sequence(64, 64, 1, seed=2)
    init = port.dxt1_fit(frames[0])
    blocks, motion, unique = port.reencode(frames[0], True, 2, 50, init, None)
    bad = blocks.copy()
    bad[37] ^= np.uint64(1 << 40)
    assert port.check_blocks(frames[0], True, 2, 50, init, bad, None, motion, np.arange(blocks.size)) >= 1


def test_endpoint_planes_padding_extension():
    """Non-multiple-of-64 planes: edge replication up to the next multiple of 64 (extension)."""
    rng = np.random.default_rng(1)
    bw, bh = 100, 70
    blocks = rng.integers(0, 2**63, size=bw * bh, dtype=np.uint64)
    got = port.endpoint_planes(blocks, bw, bh)
    assert got.shape == (6, 128, 128)
    padded = np.empty((128, 128), dtype=np.uint64)
    b2 = blocks.reshape(bh, bw)
    padded[:] = b2[np.minimum(np.arange(128), bh - 1)[:, None], np.minimum(np.arange(128), bw - 1)[None, :]]
    assert np.array_equal(got, port.endpoint_planes(padded.reshape(-1), 128, 128))


# ---- decoder side, pinned on outputs of the reference itself --------------------
def test_arith_decoder_inverts_reference_encodings():
    g = load("arith")
    for k in [k[4:] for k in g.files if k.startswith("sym_")]:
        sym = g["sym_" + k]
        assert np.array_equal(port.arith_decode(g["enc_" + k].tobytes(), sym.size), sym), k


def test_inverse_planes_recover_reference_endpoints():
    """The reference's own symbol planes -> the endpoints of the reference's own final blocks."""
    g = load("seq_256x256_sa8")
    for i in range(2):
        ep1, ep2 = port.inverse_planes(g[f"planes_{i}"], 64, 64)
        final = g[f"final_{i}"]
        assert np.array_equal(ep1, (final & np.uint64(0xFFFF)).astype(np.uint16))
        assert np.array_equal(ep2, ((final >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint16))


def test_reference_payload_decodes_to_reference_symbols():
    g = load("seq_256x256_sa8")
    import struct
    for i in range(2):
        payload = g[f"payload_{i}"].tobytes()
        off = 4
        want = [g[f"motion_{i}"], g[f"planes_{i}"][0], np.concatenate([g[f"planes_{i}"][1], g[f"planes_{i}"][2]]),
                g[f"planes_{i}"][3], np.concatenate([g[f"planes_{i}"][4], g[f"planes_{i}"][5]])]
        for s in range(5):
            (n,) = struct.unpack_from("<I", payload, off)
            off += 4
            assert np.array_equal(port.arith_decode(payload[off:off + n], want[s].size), want[s].reshape(-1)), (i, s)
            off += n
        assert off == len(payload)


def test_reference_stream_decodes_to_reference_blocks():
    """The stream CompressMultiUnique wrote (fixture) -> the blocks the encoder produced."""
    g = load("stream_256x256_sa4_gop2")
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    decoded = port.decode_stream(g["stream"].tobytes())
    prev = None
    for i in range(n // gop * gop):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        assert np.array_equal(decoded[i], blocks), f"frame {i}"
        prev = blocks


def test_decoded_pictures_match_the_reference_decoder():
    """Fixture from the reference's own decoder functions (gen_golden.decode_fixture): its decoded
    blocks equal the restated stream decode, and its DecompressedImage equals decode_rgb."""
    g = load("stream_256x256_sa4_gop2")
    d = load("decode_256x256_sa4_gop2")
    w, h = int(g["params"][0]), int(g["params"][1])
    decoded = port.decode_stream(g["stream"].tobytes())
    assert np.array_equal(decoded, d["blocks"])
    for i in range(decoded.shape[0]):
        rgb = port.decode_rgb(decoded[i], w, h)
        assert sha(rgb) == str(d["rgb_sha"][i]), f"frame {i}"
    assert np.array_equal(port.decode_rgb(decoded[-1], w, h)[:16], d["rgb_last_rows"])


def test_inter_pixel_search_port_equals_reference_fixture():
    """mptc_oracle_inter_pixel_search against the fixture produced by the reference's own CompressedBlock
    methods (DXTImage::InterPixelSearch, dxt_image.cpp:776-832, with the undefined behaviour of
    Get4X4InterpolationBlock removed: tests/golden/gen_golden.py inter_pixel_fixture)."""
    from mptc_b200.synth import make_sequence
    g = load("inter_pixel_search")
    for name in ("a", "b"):
        w, h, seed, sa = [int(x) for x in g[f"{name}_params"]]
        fr = make_sequence(w, h, 2, seed=seed)
        assert np.array_equal(port.dxt1_fit(fr[1]), g[f"{name}_cur"])
        got = port.inter_pixel_search(fr[1], sa, g[f"{name}_cur"], g[f"{name}_prev"])
        for k in ("min_err", "motion", "index", "reassigned"):
            assert np.array_equal(got[k], g[f"{name}_{k}"]), (name, k)
        n = port.ips_pattern(sa)
        assert n.shape[0] == 1 + 4 * sa * (sa - 1) and tuple(n[0]) == (0, 0) and np.abs(n).max() == sa - 1
