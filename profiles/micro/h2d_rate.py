"""Host -> device rate of the box for the benchmark's frames: one 373 MB copy, and 60 copies of 6.2 MB (pinned memory,
one stream).  usage: python profiles/micro/h2d_rate.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402

W, H, N = 1920, 1080, 60
pin = capi.PinnedArray((N, H, W, 3), np.uint8)
pin.array[:] = 7
host = torch.from_numpy(pin.array)
dev = torch.empty((N, H, W, 3), dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
for label, chunks in (("one copy of 373 MB", 1), ("60 copies of 6.2 MB", N), ("240 copies of 1.6 MB", 4 * N)):
    best = 1e9
    hv, dv = host.view(chunks, -1), dev.view(chunks, -1)
    for _ in range(5):
        torch.cuda.synchronize()
        t = time.perf_counter()
        with torch.cuda.stream(s):
            for i in range(chunks):
                dv[i].copy_(hv[i], non_blocking=True)
        s.synchronize()
        best = min(best, time.perf_counter() - t)
    print(f"{label:22s} {best * 1e3:6.2f} ms  {host.numel() / best / 1e9:5.1f} GB/s")
