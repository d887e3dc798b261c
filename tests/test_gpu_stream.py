"""GPU: whole encoder (CUDA hot path + host arithmetic coder) against the stream the reference's
CompressMultiUnique wrote for the same frames -- byte identical."""
import numpy as np
import pytest

from golden_util import load, sha
from mptc_b200 import capi
from mptc_b200.synth import make_sequence

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fixture", ["stream_256x256_sa4_gop2", "stream_512x256_sa16_gop4"])
@pytest.mark.parametrize("threads", [1, 8])
def test_encode_stream_equals_reference(ctx, threads, fixture):
    g = load(fixture)
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = make_sequence(w, h, n, seed=seed)
    assert sha(frames) == str(g["frames_sha"])
    stream, st = capi.encode_stream(ctx, frames, sa, thr, gop, threads)
    assert stream == g["stream"].tobytes()
    assert st.gpu_ms > 0
