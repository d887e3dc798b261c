"""CPU, build container only: the oracle restatement against the compiled reference
(oracle/_ref/libmptc_ref.so).  Skipped where /root/reference (hence _ref) does not exist --
the committed golden fixtures cover that case."""
import numpy as np
import pytest

from mptc_b200.synth import make_sequence
from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference)")


def test_raw_constructor_equals_png_constructor():
    assert ref.selfcheck_png(make_sequence(96, 64, 1, seed=4)[0]) == 0


@pytest.mark.parametrize("w,h,n,seed,sa,thr,gop", [(96, 64, 3, 1, 3, 50, 3), (160, 128, 2, 2, 6, 5, 2),
                                                   (64, 64, 2, 3, 1, 200, 1), (128, 128, 2, 9, 16, 50, 2)])
def test_port_equals_reference(w, h, n, seed, sa, thr, gop):
    frames = make_sequence(w, h, n, seed=seed)
    seq = ref.encode_sequence(frames, sa, thr, gop)
    prev = None
    for i, fr in enumerate(seq):
        init = port.dxt1_fit(frames[i])
        assert np.array_equal(init, fr.initial_blocks)
        blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
        assert np.array_equal(blocks, fr.blocks())
        assert np.array_equal(motion, fr.motion())
        assert np.array_equal(unique, fr.unique())
        assert abs(port.psnr(frames[i], blocks) - fr.psnr_physical()) < 1e-9
        prev = blocks


def test_random_noise_and_flat_content():
    rng = np.random.default_rng(8)
    frames = rng.integers(0, 256, size=(2, 64, 64, 3), dtype=np.uint8)
    frames[:, 16:48, 16:48] = 77  # flat: den == 0 path (dxt_image.cpp:320)
    seq = ref.encode_sequence(frames, 4, 50, 2)
    prev = None
    for i, fr in enumerate(seq):
        init = port.dxt1_fit(frames[i])
        blocks, motion, unique = port.reencode(frames[i], i == 0, 4, 50, init, prev)
        assert np.array_equal(init, fr.initial_blocks)
        assert np.array_equal(blocks, fr.blocks()) and np.array_equal(motion, fr.motion())
        prev = blocks


def test_planes_and_payload():
    frames = make_sequence(256, 256, 1, seed=6)
    fr = ref.encode_sequence(frames, 4, 50, 1)[0]
    payload = fr.entropy_payload()
    nu, planes, motion, sizes = fr.payload_planes(payload)
    mine = port.endpoint_planes(fr.blocks(), 64, 64).reshape(6, -1)
    assert np.array_equal(mine, planes)
    assert port.arith_encode(planes[0]) == ref.arith_encode(planes[0])
    assert len(port.arith_encode(motion)) == int(sizes[0])


def test_decoder_restatement_equals_reference_decoder():
    """Encode with the product's host coder from reference results, decode with the reference's own
    decoder functions and with the restatement: same blocks, same pictures."""
    from mptc_b200 import capi
    w, h, n, sa, thr, gop = 256, 256, 4, 6, 30, 2
    frames = make_sequence(w, h, n, seed=31)
    seq = ref.encode_sequence(frames, sa, thr, gop)
    nb = (w // 4) * (h // 4)
    motion = np.stack([fr.motion() for fr in seq])
    n_unique = np.array([fr.unique().size for fr in seq], dtype=np.uint32)
    unique = np.zeros((n, nb), dtype=np.uint32)
    for i, fr in enumerate(seq):
        unique[i, : n_unique[i]] = fr.unique()
    planes = np.stack([port.endpoint_planes(fr.blocks(), w // 4, h // 4) for fr in seq])
    stream, _ = capi.assemble_stream(w, h, sa, thr, gop, motion, unique, n_unique, planes)
    blocks, rgb = ref.decode_stream(stream)
    assert np.array_equal(blocks, np.stack([fr.blocks() for fr in seq]))
    assert np.array_equal(port.decode_stream(stream), blocks)
    for i in range(n):
        assert np.array_equal(port.decode_rgb(blocks[i], w, h), rgb[i])


def test_inter_pixel_search_port_vs_reference_methods():
    """The C restatement against the loop of DXTImage::InterPixelSearch run over the reference's own
    CompressedBlock methods (defined variant, oracle/ref_wrap.cpp)."""
    from mptc_b200.synth import make_sequence
    for w, h, sa, seed in [(64, 48, 2, 1), (128, 64, 8, 3), (64, 64, 20, 9)]:
        fr = make_sequence(w, h, 2, seed=seed)
        a = ref.RefFrame(fr[0], True, sa, 50)
        a.reencode(None)
        b = ref.RefFrame(fr[1], False, sa, 50)
        want = b.inter_pixel_search(a, sa)
        got = port.inter_pixel_search(fr[1], sa, b.blocks(), a.blocks())
        for k in want:
            assert np.array_equal(got[k], want[k]), (w, h, sa, k)
