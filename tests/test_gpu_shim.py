"""GPU: the reference-side shim of INTEGRATION.md, compiled for real.

oracle/_ref/libmptc_shim.so (`make -C oracle shim`) = integration/dxt_image_gpu.h + the UNMODIFIED
reference sources, linked with libmptc_b200.so.  It drives the reference's frame loop
(CompressMultiUnique, codec/codec.cpp:1383-1509) with MakeReencodedFrame -- the GPU path through the C
ABI -- in place of `new DXTImage(file, ...)` + `Reencode(prev, -1)`, then hands the GPU-filled DXTImage
to the reference's OWN EntropyEncode (codec.cpp:1115-1158) and Get8BitPalette (dxt_image.h:117-122).
The payload bytes must equal the golden payload of the reference's CPU path (tests/golden/gen_golden.py).
Also here: two host threads, each with its own context on the same GPU (mptc_gpu.h: one context per
(host thread, GPU)) -- first launches race on the launchers' per-device caches unless those are guarded."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from golden_util import load, sha
from mptc_b200 import capi
from mptc_b200.synth import make_sequence

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "libmptc_shim.so")


def _shim():
    if not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/libmptc_shim.so not built (make -C oracle shim needs /root/reference)")
    capi.load()                                  # libmptc_b200.so first: the shim links against it
    L = C.CDLL(SHIM)
    vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
    L.mptc_shim_encode_frames.argtypes = [ci, vp, ci, ci, ci, ci, ci, ci, vp, sz, vp, vp, sz, vp, vp]
    return L


def test_reference_entropy_encode_on_gpu_filled_dxtimage_gives_the_golden_payload():
    g = load("seq_256x256_sa8")
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    assert sha(frames) == str(g["frames_sha"])
    nb = (w // 4) * (h // 4)
    payload = np.zeros(n * (16 * nb + (1 << 20)), np.uint8)
    palette = np.zeros(n * nb * 4, np.uint8)
    pay_sz, pal_sz = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    blocks = np.zeros((n, nb), np.uint64)
    r = _shim().mptc_shim_encode_frames(0, frames.ctypes.data, n, w, h, sa, thr, gop, payload.ctypes.data, payload.size,
                                        pay_sz.ctypes.data, palette.ctypes.data, palette.size, pal_sz.ctypes.data,
                                        blocks.ctypes.data)
    assert r == 0
    po = qo = 0
    for i in range(n):
        assert np.array_equal(blocks[i], g[f"final_{i}"]), f"frame {i}: blocks"
        got = payload[po:po + int(pay_sz[i])]
        assert got.tobytes() == g[f"payload_{i}"].tobytes(), f"frame {i}: EntropyEncode payload differs from the CPU path's"
        pal = palette[qo:qo + int(pal_sz[i])]
        assert pal.tobytes() == g[f"unique_{i}"].astype("<u4").tobytes(), f"frame {i}: Get8BitPalette"
        po += int(pay_sz[i])
        qo += int(pal_sz[i])


def test_two_host_threads_two_contexts_one_gpu():
    """Fresh contexts in two threads, first encode of different search areas at the same time."""
    from oracle import port
    cases = [(128, 96, 3, 11, 4, 10, 3), (64, 64, 3, 5, 2, 50, 3), (192, 64, 2, 9, 16, 50, 2), (96, 96, 2, 2, 8, 0, 1)]
    want = []
    for w, h, n, seed, sa, thr, gop in cases:
        frames = make_sequence(w, h, n, seed=seed)
        prev, res = None, []
        for i in range(n):
            init = port.dxt1_fit(frames[i])
            blocks, motion, unique = port.reencode(frames[i], i % gop == 0, sa, thr, init, prev)
            res.append((blocks, motion, unique))
            prev = blocks
        want.append((frames, res))
    errors = []
    start = threading.Barrier(2)

    def worker(which):
        try:
            ctx = capi.Context(0)
            start.wait()
            for rep in range(3):
                for ci in which:
                    w, h, n, seed, sa, thr, gop = cases[ci]
                    frames, res = want[ci]
                    out = ctx.encode_sequence(frames, sa, thr, gop)
                    for i in range(n):
                        blocks, motion, unique = res[i]
                        assert np.array_equal(out["blocks"][i], blocks), (ci, i)
                        assert np.array_equal(out["motion"][i], motion), (ci, i)
                        assert np.array_equal(out["unique"][i, : out["n_unique"][i]], unique), (ci, i)
            ctx.close()
        except BaseException as e:   # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(w,)) for w in ((0, 2, 1, 3), (2, 3, 0, 1))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
