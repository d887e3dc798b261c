"""Host arithmetic decoder, one thread: Msym/s for 1 / 2 / 4 / 8 interleaved streams (mptc_arith_decode_multi) on
symbols shaped like the endpoint planes (normal around 128) and like motion bytes.  usage: python profiles/micro/decoder_bench.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402

rng = np.random.default_rng(0)
N = 2_000_000
kinds = {"planes (sigma 3)": lambda: np.clip(rng.normal(128, 3, N), 0, 255).astype(np.uint8),
         "planes (sigma 12)": lambda: np.clip(rng.normal(128, 12, N), 0, 255).astype(np.uint8),
         "uniform bytes": lambda: rng.integers(0, 256, N, dtype=np.uint8)}
for name, gen in kinds.items():
    syms = [gen() for _ in range(8)]
    codes = [capi.arith_encode(s) for s in syms]
    res = []
    for k in (1, 2, 4, 8):
        best = 0.0
        for _ in range(4):
            t = time.perf_counter()
            outs = capi.arith_decode_multi(codes[:k], [N] * k)
            best = max(best, k * N / (time.perf_counter() - t) / 1e6)
        assert all(np.array_equal(o, s) for o, s in zip(outs, syms[:k]))
        res.append(f"{k} stream(s) {best:6.1f}")
    print(f"{name:18s} " + "  ".join(res) + "  Msym/s per thread")
