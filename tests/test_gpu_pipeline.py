"""GPU: the frame-step pipelined schedule (GOP lanes, per-frame H2D / D2H, asynchronous entry
point) and the overlapped host arithmetic coder give exactly the results of the plain path:
bit-exact against the oracle, identical for every lane count, and a byte-identical stream."""
import numpy as np
import pytest

from mptc_b200 import capi
from mptc_b200.synth import make_sequence
from oracle import port

pytestmark = pytest.mark.gpu


def oracle_sequence(frames, sa, thr, gop):
    prev, res = None, []
    for i, rgb in enumerate(frames):
        init = port.dxt1_fit(rgb)
        blocks, motion, unique = port.reencode(rgb, i % gop == 0, sa, thr, init, prev)
        res.append((blocks, motion, unique))
        prev = blocks
    return res


@pytest.mark.parametrize("lanes", [1, 2, 3, 5])
def test_lanes_and_ragged_last_gop_bit_exact(ctx, lanes):
    """11 frames, gop 3: three full GOPs + a ragged one of 2 frames, split over 1..4 lanes."""
    w, h, n, sa, thr, gop = 192, 128, 11, 4, 30, 3
    frames = make_sequence(w, h, n, seed=21)
    ref = oracle_sequence(frames, sa, thr, gop)
    ctx.set_schedule(lanes, 0, 0)
    try:
        out = ctx.encode_sequence(frames, sa, thr, gop)
    finally:
        ctx.set_schedule(0, 0, 0)
    for i in range(n):
        blocks, motion, unique = ref[i]
        assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
        assert np.array_equal(out["motion"][i], motion), f"frame {i}"
        nu = int(out["n_unique"][i])
        assert nu == unique.size and np.array_equal(out["unique"][i, :nu], unique)
        assert np.array_equal(out["planes"][i], port.endpoint_planes(blocks, w // 4, h // 4)), f"planes {i}"


def test_async_wait_frame_matches_sync(ctx):
    """mptc_gpu_encode_sequence_async + mptc_gpu_wait_frame: a frame's host results are complete
    when its wait returns (frames are consumed here in arrival order, k-major)."""
    w, h, n, sa, thr, gop = 512, 256, 12, 8, 50, 4
    frames_p = capi.PinnedArray((n, h, w, 3), np.uint8)
    frames_p.array[:] = make_sequence(w, h, n, seed=5)
    frames = frames_p.array
    sync = ctx.encode_sequence(frames, sa, thr, gop)
    nb = (w // 4) * (h // 4)
    pbw, pbh = (w // 4 + 63) // 64 * 64, (h // 4 + 63) // 64 * 64
    pins = {"blocks": capi.PinnedArray((n, nb), np.uint64), "motion": capi.PinnedArray((n, 2 * nb), np.uint8),
            "unique": capi.PinnedArray((n, nb), np.uint32), "n_unique": capi.PinnedArray((n,), np.uint32),
            "planes": capi.PinnedArray((n, 6, pbh, pbw), np.uint8)}
    out = {k: v.array for k, v in pins.items()}
    for a in out.values():
        a[...] = 0
    ctx.encode_sequence(frames, sa, thr, gop, out=out, wait=False)
    for k in range(gop):
        for g in range(n // gop):
            f = g * gop + k
            ctx.wait_frame(f)
            assert np.array_equal(out["blocks"][f], sync["blocks"][f]), f"frame {f}"
            assert np.array_equal(out["motion"][f], sync["motion"][f]), f"frame {f}"
            assert out["n_unique"][f] == sync["n_unique"][f]
            assert np.array_equal(out["planes"][f], sync["planes"][f]), f"frame {f}"
    ctx.wait()
    with pytest.raises(capi.MptcError):
        ctx.wait_frame(n)


@pytest.mark.parametrize("threads", [1, 6])
def test_overlapped_stream_equals_two_phase_assembly(ctx, threads):
    """mptc_encode_stream (arithmetic coder overlapped with the GPU) == mptc_assemble_stream over
    the results of a plain encode; 10 frames with gop 4 leave a trailing partial group unwritten."""
    w, h, n, sa, thr, gop = 256, 256, 10, 8, 50, 4
    frames = make_sequence(w, h, n, seed=9)
    out = ctx.encode_sequence(frames, sa, thr, gop)
    two_phase, st2 = capi.assemble_stream(w, h, sa, thr, gop, out["motion"], out["unique"], out["n_unique"],
                                          out["planes"], threads=2)
    stream, st = capi.encode_stream(ctx, frames, sa, thr, gop, threads)
    assert stream == two_phase
    assert st.n_groups == 2 and st.total_ms >= st.entropy_ms > 0 and st.gpu_ms > 0
    assert (st.max_comp_motion, st.max_comp_ep_y, st.max_comp_ep_c) == (st2.max_comp_motion, st2.max_comp_ep_y, st2.max_comp_ep_c)


@pytest.mark.parametrize("max_pct", [0, 100])
@pytest.mark.parametrize("sa,thr", [(4, 0), (16, 0), (2, 50), (20, 5)])
def test_leftover_paths_bit_exact(max_pct, sa, thr, monkeypatch):
    """Inter frames with MANY leftover blocks (thr 0 / tiny windows), forced through the per-block
    kernel K3s (max_pct 100: never hands a frame to the wavefront; sa 20 takes its position-by-
    position path) and through the row wavefront K3 (max_pct 0) -- both equal the oracle."""
    monkeypatch.setenv("MPTC_SPARSE_MAX_PCT", str(max_pct))
    c = capi.Context(0)
    try:
        w, h, n, gop = 256, 192, 6, 3
        frames = make_sequence(w, h, n, seed=33)
        ref = oracle_sequence(frames, sa, thr, gop)
        out = c.encode_sequence(frames, sa, thr, gop)
        left = 0
        for i in range(n):
            blocks, motion, unique = ref[i]
            assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
            assert np.array_equal(out["motion"][i], motion), f"frame {i}"
            assert int(out["n_unique"][i]) == unique.size
            if i % gop:
                m = motion.reshape(-1, 2)
                left += int(((m[:, 0] & 0x80) == 0).sum() + ((m[:, 0] == 255) & (m[:, 1] == 255)).sum())
        assert left > 0, "the case must exercise the leftover path"
    finally:
        c.close()


def test_two_devices_in_one_process():
    """One context per GPU inside one process: per-device launch
    configuration, same results on every device.  Needs two GPUs; skipped otherwise."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h, n, sa, thr, gop = 256, 128, 4, 16, 50, 2
    frames = make_sequence(w, h, n, seed=13)
    ref = oracle_sequence(frames, sa, thr, gop)
    ctxs = [capi.Context(d) for d in (1, 0)]
    try:
        for c in ctxs:
            out = c.encode_sequence(frames, sa, thr, gop)
            for i in range(n):
                assert np.array_equal(out["blocks"][i], ref[i][0]), f"device {c.device} frame {i}"
            dec = c.decode_sequence(out["motion"], out["unique"], out["n_unique"], out["planes"], w, h, sa, gop)
            assert np.array_equal(dec, out["blocks"])
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("lanes,wave_rows,gop,w", [(1, 1, 1, 320), (1, 1, 2, 512), (2, 1, 1, 320), (4, 2, 1, 272)])
def test_fewer_wavefront_ctas_than_gops(ctx, lanes, wave_rows, gop, w):
    """rows wider than one group are shared by several CTAs; with a grid no larger than
    the number of GOPs every resident CTA used to hold a first-part ticket and nobody could pick up the
    partner's.  The parts of a row now hold adjacent tickets.  Here: up to 10 GOPs per launch, one or two
    CTAs per frame, rows of 3 - 4 groups."""
    h, n, sa, thr = 48, 10, 4, 50
    frames = make_sequence(w, h, n, seed=21)
    ref = oracle_sequence(frames, sa, thr, gop)
    ctx.set_schedule(lanes, wave_rows, wave_rows)
    try:
        out = ctx.encode_sequence(frames, sa, thr, gop)
    finally:
        ctx.set_schedule(0, 0, 0)
    for i in range(n):
        blocks, motion, unique = ref[i]
        assert np.array_equal(out["blocks"][i], blocks), f"frame {i}"
        assert np.array_equal(out["motion"][i], motion), f"frame {i}"
