#!/usr/bin/env python
"""Full-size fixtures of the BENCHMARKED configurations, generated from the UNMODIFIED reference
(oracle/_ref/libmptc_ref.so = /root/reference compiled by oracle/Makefile).

  python tests/golden/gen_golden_full.py [case ...]      (only where /root/reference exists)

The reference's frame loop (CompressMultiUnique, codec.cpp:1383-1509: DXTImage ctor ->
Reencode(prev, -1), dxt_image.cpp:868-957) is run over whole GOPs of the sequences bench.py times
(mptc_b200.synth.make_frame, seed 1234, frames numbered from 0):

  full_1080p60_sa16_gop15   1920x1080 x 60 frames = the whole BASELINE configs[1] batch of rank 0
  full_4k15_sa16_gop15      3840x2160 x 15 frames = the first GOP of configs[2] / configs[4]

GOPs are independent (an intra frame ignores its predecessor, dxt_image.cpp:885), so each GOP is a
separate process.  Stored per frame: SHA-256 of the initial blocks, final blocks, motion bytes and
unique palette, the unique count, and CRC-32 of every block ROW of the final blocks and of the
motion bytes (so a mismatch can be located without the 1 MB / frame of raw blocks).  The reference
costs 15-30 s per 1080p frame per core at sa 16: the 1080p case is about 8 minutes on 4 cores,
the 4K GOP about half an hour on one.
"""
import hashlib
import multiprocessing as mp
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name: (w, h, n_frames, seed, search_area, err_threshold, gop)
CASES = {
    "full_1080p60_sa16_gop15": (1920, 1080, 60, 1234, 16, 50, 15),
    "full_4k15_sa16_gop15": (3840, 2160, 15, 1234, 16, 50, 15),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def row_crcs(a: np.ndarray, rows: int) -> np.ndarray:
    a = np.ascontiguousarray(a).reshape(rows, -1)
    return np.array([zlib.crc32(r.tobytes()) for r in a], dtype=np.uint32)


def run_gop(job):
    w, h, f0, n, seed, sa, thr = job
    from mptc_b200.synth import make_frame
    from oracle import ref
    bh = h // 4
    prev, out = None, []
    for k in range(n):
        t0 = time.time()
        fr = ref.RefFrame(make_frame(w, h, f0 + k, seed), k == 0, sa, thr)
        init = fr.blocks()
        fr.reencode(prev)
        fin, mo, un = fr.blocks(), fr.motion(), fr.unique()
        out.append({"hashes": [sha(init), sha(fin), sha(mo), sha(un)], "n_unique": un.size,
                    "rows_final": row_crcs(fin, bh), "rows_motion": row_crcs(mo, bh),
                    "n_inter": int(np.count_nonzero(((mo[0::2] & 0x80) != 0) & (mo[0::2] != 255)))})
        prev = fr
        print(f"  frame {f0 + k} of {w}x{h}: {time.time() - t0:.1f} s, {un.size} unique", flush=True)
    return f0, out


def run_case(name):
    w, h, n, seed, sa, thr, gop = CASES[name]
    jobs = [(w, h, f0, min(gop, n - f0), seed, sa, thr) for f0 in range(0, n, gop)]
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = dict(pool.map(run_gop, jobs))
    frames = [fr for f0 in sorted(res) for fr in res[f0]]
    from mptc_b200.synth import make_frame
    h_in = hashlib.sha256()                      # the input the fixture belongs to
    for f in range(n):
        h_in.update(make_frame(w, h, f, seed).tobytes())
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        params=np.array([w, h, n, seed, sa, thr, gop], dtype=np.int64),
        frames_sha=np.array(h_in.hexdigest()),
        hashes=np.array([f["hashes"] for f in frames]),
        n_unique=np.array([f["n_unique"] for f in frames], dtype=np.uint32),
        n_inter=np.array([f["n_inter"] for f in frames], dtype=np.uint32),
        rows_final=np.stack([f["rows_final"] for f in frames]),
        rows_motion=np.stack([f["rows_motion"] for f in frames]))
    print("wrote", name, flush=True)


if __name__ == "__main__":
    from oracle import ref
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)
