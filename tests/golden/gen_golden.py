#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libmptc_ref.so, built
by `make -C oracle ref` from /root/reference).  The reference ships no golden vectors of its
own, so these fixtures -- outputs of the reference itself, run in the build
container -- are what pins the oracle and the CUDA path.  Re-run only where /root/reference
exists:   python tests/golden/gen_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mptc_b200.synth import make_sequence  # noqa: E402
from oracle import ref  # noqa: E402

# name: (w, h, n_frames, seed, search_area, err_threshold, gop, store_arrays)
CASES = {
    "seq_64x64_sa2": (64, 64, 3, 5, 2, 50, 3, True),
    "seq_128x96_sa4_thr10": (128, 96, 3, 11, 4, 10, 3, True),
    "seq_256x256_sa8": (256, 256, 2, 1234, 8, 50, 2, True),       # multiple of 256: planes + payload defined
    "seq_256x256_sa16_hash": (256, 256, 8, 1234, 16, 50, 4, False),  # BASELINE configs[0], hashes only
    "seq_320x192_sa4_thr0_hash": (320, 192, 6, 3, 4, 0, 3, False),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(w, h, n, seed, sa, thr, gop, store):
    frames = make_sequence(w, h, n, seed=seed)
    seq = ref.encode_sequence(frames, sa, thr, gop)
    out = {"params": np.array([w, h, n, seed, sa, thr, gop], dtype=np.int64), "frames_sha": np.array(sha(frames))}
    hashes = []
    for i, fr in enumerate(seq):
        init, fin, mo, un = fr.initial_blocks, fr.blocks(), fr.motion(), fr.unique()
        hashes.append([sha(init), sha(fin), sha(mo), sha(un)])
        if store:
            out[f"init_{i}"] = init
            out[f"final_{i}"] = fin
            out[f"motion_{i}"] = mo
            out[f"unique_{i}"] = un
        out[f"psnr_physical_{i}"] = np.array(fr.psnr_physical())
        if w % 256 == 0 and h % 256 == 0:
            payload = fr.entropy_payload()
            nu, planes, motion, sizes = fr.payload_planes(payload)
            assert nu == un.size and np.array_equal(motion, mo)
            out[f"sizes_{i}"] = sizes
            hashes[-1] += [sha(planes), hashlib.sha256(payload).hexdigest()]
            if store:
                out[f"planes_{i}"] = planes
                out[f"payload_{i}"] = np.frombuffer(payload, dtype=np.uint8)
    out["hashes"] = np.array(hashes)
    return out


def reference_stream(frames, sa, thr, gop):
    """Runs the reference's own CompressMultiUnique (codec.cpp:1307) on a PNG directory in a
    fresh process (its max_* header fields are process-global, codec.cpp:73-77)."""
    import subprocess
    import tempfile
    from PIL import Image
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "in"))
        for i, fr in enumerate(frames):
            Image.fromarray(fr, "RGB").save(os.path.join(d, "in", f"{i:05d}.png"))
        out = os.path.join(d, "out.mptc")
        code = ("import sys; sys.path.insert(0, %r); from oracle import ref; "
                "ref.lib().mptc_ref_compress_multi_unique(%r.encode(), %r.encode(), %d, %d, %d, %d)"
                % (ROOT, os.path.join(d, "in"), out, sa, thr, gop, gop))
        subprocess.check_call([sys.executable, "-c", code], cwd=d, stdout=subprocess.DEVNULL)
        return open(out, "rb").read()


def decode_fixture():
    """The reference's own decoder (mptc_ref_decode_stream: the loop of DecompressMultiUnique over
    the reference's EntropyDecode / ReconstructDXTData / ReconstructEndPoints / DecompressedImage)
    run on the stream fixture: decoded blocks of every frame, hashes of the decoded pictures and
    the first 16 pixel rows of the last one."""
    g = np.load(os.path.join(HERE, "stream_256x256_sa4_gop2.npz"))
    blocks, rgb = ref.decode_stream(g["stream"].tobytes())
    np.savez_compressed(os.path.join(HERE, "decode_256x256_sa4_gop2.npz"), blocks=blocks,
                        rgb_sha=np.array([sha(f) for f in rgb]), rgb_last_rows=rgb[-1, :16].copy())
    print("wrote decode fixture", blocks.shape)


STREAM_CASES = {"stream_256x256_sa4_gop2": (256, 256, 4, 77, 4, 50, 2), "stream_512x256_sa16_gop4": (512, 256, 8, 41, 16, 50, 4)}


def stream_fixtures(only=None):
    """The bytes the reference's own CompressMultiUnique (codec.cpp:1307) writes for a PNG directory."""
    for name, (w, h, n, seed, sa, thr, gop) in STREAM_CASES.items():
        if only and name != only:
            continue
        frames = make_sequence(w, h, n, seed=seed)
        stream = reference_stream(frames, sa, thr, gop)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), params=np.array([w, h, n, seed, sa, thr, gop]),
                            frames_sha=np.array(sha(frames)), stream=np.frombuffer(stream, dtype=np.uint8))
        print("wrote", name, len(stream), "bytes")


def inter_pixel_fixture():
    """DXTImage::InterPixelSearch (dxt_image.cpp:776-832), every block of the second frame of two small
    sequences against the first frame's final blocks: the loop run over the reference's own
    CompressedBlock methods with the candidate word built without the undefined behaviour of
    Get4X4InterpolationBlock (oracle/ref_wrap.cpp, mptc_ref_inter_pixel_search_defined)."""
    out = {}
    for name, (w, h, seed, sa) in {"a": (128, 96, 11, 4), "b": (96, 64, 5, 16)}.items():
        fr = make_sequence(w, h, 2, seed=seed)
        a = ref.RefFrame(fr[0], True, sa, 50)
        a.reencode(None)
        b = ref.RefFrame(fr[1], False, sa, 50)
        r = b.inter_pixel_search(a, sa)
        ub = b.inter_pixel_search(a, sa, defined=False)
        out[f"{name}_params"] = np.array([w, h, seed, sa], dtype=np.int64)
        out[f"{name}_prev"] = a.blocks()
        out[f"{name}_cur"] = b.blocks()
        for k, v in r.items():
            out[f"{name}_{k}"] = v
        out[f"{name}_ub_agrees"] = np.array(int((ub["index"] == r["index"]).sum()))   # for the record
    np.savez_compressed(os.path.join(HERE, "inter_pixel_search.npz"), **out)
    print("wrote inter_pixel_search")


def main():
    if "--decode-only" in sys.argv:
        return decode_fixture()
    if "--inter-pixel-only" in sys.argv:
        return inter_pixel_fixture()
    if "--stream2-only" in sys.argv:
        return stream_fixtures("stream_512x256_sa16_gop4")
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    assert ref.selfcheck_png(make_sequence(64, 64, 1)[0]) == 0
    for name, cfg in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(*cfg))
        print("wrote", name)
    # arithmetic coder known answers
    rng = np.random.default_rng(99)
    streams = {
        "empty": np.zeros(0, dtype=np.uint8),
        "one": np.array([200], dtype=np.uint8),
        "zeros": np.zeros(3000, dtype=np.uint8),
        "skewed": np.clip(rng.normal(128, 6, 20000), 0, 255).astype(np.uint8),
        "uniform": rng.integers(0, 256, 5000, dtype=np.uint8),
        "ff": np.full(700, 255, dtype=np.uint8),
    }
    out = {}
    for k, s in streams.items():
        out["sym_" + k] = s
        out["enc_" + k] = np.frombuffer(ref.arith_encode(s), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "arith.npz"), **out)
    print("wrote arith")
    # whole-stream fixtures: 256x256 x 4 frames, sa 4, thr 50, gop 2 (two groups); 512x256 x 8 frames, sa 16, gop 4
    stream_fixtures()
    decode_fixture()
    inter_pixel_fixture()


if __name__ == "__main__":
    main()
