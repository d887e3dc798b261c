"""CPU: the host side above the GPU hot path (include/mptc_codec.h): arithmetic coder, frame
payload and stream assembly, bit-exact against fixtures produced by the unmodified reference
(entropy/arithmetic_codec.cpp, codec.cpp:1115-1158, CompressMultiUnique codec.cpp:1307)."""
import numpy as np
import pytest

from golden_util import load, sha
from mptc_b200 import capi
from mptc_b200.synth import make_sequence
from oracle import port


def test_arith_encode_known_answers():
    g = load("arith")
    for k in [k[4:] for k in g.files if k.startswith("sym_")]:
        assert capi.arith_encode(g["sym_" + k]) == g["enc_" + k].tobytes(), k


def test_arith_encode_matches_oracle_on_random_streams():
    rng = np.random.default_rng(5)
    for n in (1, 2, 257, 4096, 100000):
        s = np.clip(rng.normal(128, rng.uniform(1, 60), n), 0, 255).astype(np.uint8)
        assert capi.arith_encode(s) == port.arith_encode(s)


@pytest.mark.parametrize("threads", [1, 5])
def test_frame_payload_equals_reference(threads):
    g = load("seq_256x256_sa8")
    for i in range(2):
        payload, sizes = capi.frame_payload(g[f"motion_{i}"], g[f"planes_{i}"], int(g[f"unique_{i}"].size), threads)
        assert payload == g[f"payload_{i}"].tobytes()
        assert [int(x) for x in sizes] == [int(x) for x in g[f"sizes_{i}"]]


@pytest.mark.parametrize("threads,fixture", [(1, "stream_256x256_sa4_gop2"), (4, "stream_256x256_sa4_gop2"),
                                             (3, "stream_512x256_sa16_gop4")])
def test_stream_assembly_equals_reference_stream(threads, fixture):
    """Per-frame results come from the oracle here (CPU); the assembled stream must equal what the
    reference's CompressMultiUnique wrote for the same frames (two fixtures: 256x256 sa 4 gop 2 and
    512x256 sa 16 gop 4 -- the default window, four-frame groups)."""
    g = load(fixture)
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    assert sha(frames) == str(g["frames_sha"])
    nb = (w // 4) * (h // 4)
    motion = np.empty((n, 2 * nb), np.uint8)
    unique = np.zeros((n, nb), np.uint32)
    n_unique = np.zeros(n, np.uint32)
    planes = np.empty((n, 6, h // 4, w // 4), np.uint8)
    prev = None
    for i in range(n):
        init = port.dxt1_fit(frames[i])
        blocks, mo, un = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        motion[i] = mo
        unique[i, : un.size] = un
        n_unique[i] = un.size
        planes[i] = port.endpoint_planes(blocks, w // 4, h // 4)
        prev = blocks
    stream, st = capi.assemble_stream(w, h, sa, thr, gop, motion, unique, n_unique, planes, threads)
    ref = g["stream"].tobytes()
    assert len(stream) == len(ref)          # compressed size: identical, not just within 0.1 %
    assert stream == ref
    assert st.n_groups == n // gop


def test_trailing_partial_group_is_dropped_like_the_reference():
    w = h = 64
    nb = 256
    n = 3
    motion = np.full((n, 2 * nb), 255, np.uint8)
    unique = np.zeros((n, nb), np.uint32)
    n_unique = np.full(n, nb, np.uint32)
    planes = np.full((n, 6, 64, 64), 128, np.uint8)
    s3, st3 = capi.assemble_stream(w, h, 2, 50, 2, motion, unique, n_unique, planes)
    s2, st2 = capi.assemble_stream(w, h, 2, 50, 2, motion[:2], unique[:2], n_unique[:2], planes[:2])
    assert s3 == s2 and st3.n_groups == 1


# ---- decoder side --------------------------------------------------------------------------------
def test_arith_decode_inverts_reference_encodings():
    """mptc_arith_decode (EntropyDecode, codec.cpp:560-577) on the reference's own encodings."""
    g = load("arith")
    for k in [k[4:] for k in g.files if k.startswith("sym_")]:
        sym = g["sym_" + k]
        assert np.array_equal(capi.arith_decode(g["enc_" + k].tobytes(), sym.size), sym), k


def test_arith_decode_matches_oracle_and_round_trips():
    rng = np.random.default_rng(6)
    for n in (1, 2, 257, 4096, 150000):
        s = np.clip(rng.normal(128, rng.uniform(0.5, 80), n), 0, 255).astype(np.uint8)
        code = capi.arith_encode(s)
        got = capi.arith_decode(code, n)
        assert np.array_equal(got, s)
        assert np.array_equal(got, port.arith_decode(code, n))


def test_arith_decode_extreme_distributions():
    """The decoder's vector symbol search and its multi-byte renormalisation at their corners: a single repeated
    symbol (intervals stay wide: zero-byte renormalisations), all 256 symbols equally rare inside one start bucket
    (more than eight candidates per bucket: the scalar continuation), symbols 0 and 255, and short streams whose
    code ends inside the four-byte look-ahead."""
    rng = np.random.default_rng(9)
    cases = [np.full(70000, 200, np.uint8), np.zeros(3000, np.uint8), np.full(3000, 255, np.uint8),
             rng.integers(0, 256, 120000, dtype=np.uint8), np.tile(np.arange(256, dtype=np.uint8), 300),
             np.where(rng.random(90000) < 0.999, 7, rng.integers(0, 256, 90000)).astype(np.uint8),
             np.array([3], np.uint8), np.array([255, 0, 255], np.uint8)]
    for s in cases:
        code = capi.arith_encode(s)
        assert np.array_equal(capi.arith_decode(code, s.size), s)
        assert np.array_equal(port.arith_decode(code, s.size), s)


def test_arith_decode_multi_equals_single_streams():
    """mptc_arith_decode_multi: 1..8 interleaved streams, equal and ragged lengths (incl. empty), give what
    mptc_arith_decode gives stream by stream; a truncated member fails the call."""
    rng = np.random.default_rng(8)
    for k in range(1, 9):
        ns = [int(rng.integers(0, 40000)) for _ in range(k)] if k % 2 else [30000] * k
        syms = [np.clip(rng.normal(128, rng.uniform(1, 60), n), 0, 255).astype(np.uint8) for n in ns]
        codes = [capi.arith_encode(s) for s in syms]
        outs = capi.arith_decode_multi(codes, ns)
        for o, s, c in zip(outs, syms, codes):
            assert np.array_equal(o, s)
            assert np.array_equal(o, capi.arith_decode(c, s.size))
    syms = [rng.integers(0, 256, 5000, dtype=np.uint8) for _ in range(4)]
    codes = [capi.arith_encode(s) for s in syms]
    codes[2] = codes[2][: len(codes[2]) // 2]
    with pytest.raises(capi.MptcError):
        capi.arith_decode_multi(codes, [5000] * 4)
    with pytest.raises(capi.MptcError):
        capi.arith_decode_multi(codes * 3, [5000] * 12)     # more than 8 streams


def test_arith_decode_rejects_truncated_code():
    s = np.random.default_rng(7).integers(0, 256, 5000, dtype=np.uint8)
    code = capi.arith_encode(s)
    with pytest.raises(capi.MptcError):
        capi.arith_decode(code[: len(code) // 2], s.size)


def test_stream_info_reads_the_reference_header():
    g = load("stream_256x256_sa4_gop2")
    w, h, n, _seed, sa, _thr, gop = [int(x) for x in g["params"]]
    hdr = capi.stream_info(g["stream"].tobytes())
    assert (hdr.width, hdr.height, hdr.gop, hdr.search_area, hdr.n_groups, hdr.n_frames) == (w, h, gop, sa, n // gop, n)
    want = port.parse_stream(g["stream"].tobytes())["maxes"]
    assert (hdr.max_unique_bytes, hdr.max_comp_palette, hdr.max_comp_motion, hdr.max_comp_ep_y, hdr.max_comp_ep_c) == tuple(want)
    with pytest.raises(capi.MptcError):
        capi.stream_info(g["stream"].tobytes()[:20])


def test_stream_info_rejects_headers_that_promise_more_than_the_stream_holds():
    """A crafted 34-byte header (2^24 groups x gop 255) must be reported as corrupt before anything
    is sized by it (mptc_decode_stream allocates per group / per frame)."""
    import struct
    hdr = struct.pack("<IIBBI", 256, 256, 255, 4, 1 << 24) + b"\0" * 20
    assert len(hdr) == 34
    with pytest.raises(capi.MptcError):
        capi.stream_info(hdr)
    g = load("stream_256x256_sa4_gop2")
    s = bytearray(g["stream"].tobytes())
    s[10:14] = struct.pack("<I", 1 << 20)           # far more groups than the stream has bytes for
    with pytest.raises(capi.MptcError):
        capi.stream_info(bytes(s))
