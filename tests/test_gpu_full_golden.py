"""GPU: the BENCHMARKED configurations against full-GOP fixtures produced by the unmodified
reference (tests/golden/gen_golden_full.py -> oracle/_ref: DXTImage ctor + Reencode over whole GOPs,
codec/dxt_image.cpp:385-436, :868-957): every block of every frame of

  * BASELINE configs[1]: 1920x1080 x 60 frames, search_area 16, err_threshold 50, GOP 15 -- the exact
    batch bench.py times on rank 0, encoded with the default schedule (4 GOP lanes, frame-index-major);
  * BASELINE configs[2]/[4]: 3840x2160, the first GOP of 15 frames.

Compared per frame: SHA-256 of the initial blocks, final blocks, motion bytes and the unique
palette (bit-exact, nothing sampled)."""
import numpy as np
import pytest

from golden_util import compare_with_full_fixture, load
from mptc_b200.synth import make_frame

pytestmark = pytest.mark.gpu


def _frames(w, h, n, seed):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as pool:
        return np.stack(list(pool.map(lambda f: make_frame(w, h, f, seed), range(n))))


def _check(g, out, first_frame=0):
    n, bad = compare_with_full_fixture(g, out["blocks"], out["motion"], out["unique"], out["n_unique"],
                                       initial=out.get("initial"), first_frame=first_frame)
    assert not bad, "\n".join(bad[:12])
    return n


def test_1080p60_gop15_every_frame_equals_the_reference(ctx):
    g = load("full_1080p60_sa16_gop15")
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = _frames(w, h, n, seed)
    # device-resident path (what bench.py's `value` times), default schedule
    ctx.seq_reserve(w, h, n)
    ctx.seq_upload(frames)
    ctx.seq_encode(0, n, sa, thr, gop)
    assert _check(g, ctx.seq_download(0, n)) == n
    # end to end from host buffers (what bench.py's `e2e` times)
    assert _check(g, ctx.encode_sequence(frames, sa, thr, gop)) == n


@pytest.mark.parametrize("lanes,wave_rows", [(1, 0), (2, 8), (4, 3)])
def test_1080p_two_gops_under_other_schedules(ctx, lanes, wave_rows):
    """The same fixture under other lane counts / wavefront widths (30 frames = 2 GOPs): the schedule
    must never change a result.  wave_rows 3 forces far fewer CTAs than rows (ticket order matters)."""
    g = load("full_1080p60_sa16_gop15")
    w, h, _n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    n = 2 * gop
    frames = _frames(w, h, n, seed)
    ctx.set_schedule(lanes, wave_rows, wave_rows)
    try:
        assert _check(g, ctx.encode_sequence(frames, sa, thr, gop)) == n
    finally:
        ctx.set_schedule(0, 0, 0)


def test_4k_first_gop_every_frame_equals_the_reference(ctx):
    import os
    from golden_util import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "full_4k15_sa16_gop15.npz")):
        pytest.skip("4K fixture not generated (tests/golden/gen_golden_full.py full_4k15_sa16_gop15)")
    g = load("full_4k15_sa16_gop15")
    w, h, n, seed, sa, thr, gop = [int(x) for x in g["params"]]
    frames = _frames(w, h, n, seed)
    assert _check(g, ctx.encode_sequence(frames, sa, thr, gop)) == n
