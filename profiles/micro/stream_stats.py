"""Where the time of mptc_encode_stream / mptc_decode_stream goes on the headline workload."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
from mptc_b200 import capi  # noqa: E402
from mptc_b200.synth import make_frame  # noqa: E402

W, H, N, SA, THR, GOP = 1920, 1080, 60, 16, 50, 15
threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
pin = capi.PinnedArray((N, H, W, 3), np.uint8)
for f in range(N):
    pin.array[f] = make_frame(W, H, f)
ctx = capi.Context(0)
buf = np.empty(pin.array.nbytes // 4 + (1 << 20), dtype=np.uint8)
for it in range(5):
    t0 = time.perf_counter()
    stream, st = capi.encode_stream(ctx, pin.array, SA, THR, GOP, threads, out=buf)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"encode_stream threads {threads}: wall {wall:.1f} ms, total {st.total_ms:.1f}, gpu {st.gpu_ms:.1f}, entropy phase {st.entropy_ms:.1f}, "
          f"assemble {st.assemble_ms:.1f}, {len(stream)} bytes")
stream = stream.tobytes()
dpin = capi.PinnedArray((N, (W // 4) * (H // 4)), np.uint64)
for it in range(3):
    t0 = time.perf_counter()
    blocks, _, ds = capi.decode_stream(ctx, stream, threads=threads, blocks_out=dpin.array)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"decode_stream threads {threads}: wall {wall:.1f} ms, total {ds.total_ms:.1f}, entropy phase {ds.entropy_ms:.1f}")
