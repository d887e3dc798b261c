#!/usr/bin/env python
"""bench.py -- MPTC encode throughput (Mpixel/s) of the B200 hot path, the driver's contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (DXT1 fit + inter/intra index-reuse search + unique
compaction + endpoint wavelet planes) over one batch: the BASELINE.json configs[1] workload,
a 1920x1080 synthetic 60-frame sequence, search_area 16, err_threshold 50, GOP 15, per GPU.
With N GPUs (torchrun, one rank per GPU) each rank encodes its own 60-frame shard of a
60*N-frame sequence: GOPs are independent, so there is no data-path collective (weak scaling).

Printed JSON (rank 0, one line):
  value     whole-job Mpixel/s with the frames already resident in HBM (device time of the
            kernels, CUDA events on the library's compute stream, max over ranks)
  e2e       the same metric through the C-ABI call with HOST (pinned) buffers: H2D of the
            frames + kernels + D2H of blocks/motion/unique/planes inside the timed region
  roofline  instruction-issue roofline of the dominant kernel: achieved =
            candidate evaluations x 360 issue slots / kernel time, peak = 148 SMs x 4
            schedulers x 32 lanes x SM clock; plus that kernel's algorithmic HBM GB/s
  cpu_baseline  the reference encoder (oracle/_ref, unmodified reference sources) on the
            box's host cores over a bounded sample of the same workload
--impl reference prints the reference arm: the reference CPU encoder alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, FRAMES, SA, THR, GOP = 1920, 1080, 60, 16, 50, 15
ALGO_SLOTS_PER_CANDIDATE = 360
SM_COUNT, SCHEDULERS, LANES = 148, 4, 32
WORKLOAD = ""
# BASELINE.json configs: [1] is the default (the headline); [2] = 4K x120 with the host coder overlapped
WORKLOADS = {"1080p60": (1920, 1080, 60), "4k120": (3840, 2160, 120), "4k600": (3840, 2160, 600)}
STRONG = {"4k600"}   # configs[4]: the frame count is that of the whole job, GOP-sharded over the GPUs
SCALING = "weak"


def set_workload(name: str, sa: int, thr: int, world: int = 1):
    global W, H, FRAMES, SA, THR, WORKLOAD, SCALING
    W, H, FRAMES = WORKLOADS[name]
    SA, THR = sa, thr
    if name in STRONG:
        total, gops = FRAMES, FRAMES // GOP
        assert gops % world == 0, f"{gops} GOPs do not split over {world} GPUs"
        FRAMES = gops // world * GOP
        SCALING = "strong"
        WORKLOAD = (f"{W}x{H} synthetic x{total} frames in all = {FRAMES} frames/GPU, search_area={SA}, "
                    f"err_threshold={THR}, gop={GOP}")
    else:
        SCALING = "weak"
        WORKLOAD = f"{W}x{H} synthetic x{FRAMES} frames/GPU, search_area={SA}, err_threshold={THR}, gop={GOP}"
METRIC = "mptc_encode_mpixel_per_s"
UNIT = "Mpixel/s"
SCAN_SLOTS_PER_POSITION = 8             # winner_update_fast: IMAD+MIN, NEG+LOP3+MIN, SHF+LOP3+MAX


def base_config(world: int) -> dict:
    """The `config` object of BOTH arms (the driver compares them key by key)."""
    rgb_mb = FRAMES * W * H * 3 // 1000000
    return {"workload": WORKLOAD, "frames_per_gpu": FRAMES, "sharding": f"gop-sharded x{world}, no collective",
            "l2": f"inputs ({rgb_mb} MB RGB per step) exceed the 126 MB L2; no explicit flush"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: an NVML polling thread (a sample every few
    milliseconds -- the timed regions of the default run are only ~0.2 s long, too short for `nvidia-smi -lms`,
    whose first sample arrives after ~0.1 s), with `nvidia-smi` as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, device_index: int):
        self.dev = device_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []       # (sm MHz, reasons bitmask)
        self.sm_max = None
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.dev < len(ids) and ids[self.dev].isdigit():
                return int(ids[self.dev])
        return self.dev

    def _one(self, nv, h):
        try:
            self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                 int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))))
        except Exception:
            pass

    def start(self):
        try:
            import pynvml as nv
            import threading
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()
            self._nv, self._h = nv, h

            def poll():
                while not self.stop_flag.is_set():
                    self._one(nv, h)
                    self.stop_flag.wait(0.004)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.dev), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._one(self._nv, self._h)      # one more while the last step's kernels have only just ended
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                bits = 0
                for _, b in self.samples:
                    bits |= b
                out.update(sm_mhz=float(np.median([x for x, _ in self.samples])), sm_max_mhz=self.sm_max,
                           samples=len(self.samples), reasons=sorted(n for m, n in self.REASONS if bits & m), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    c = [x.strip() for x in line.split(",")]
                    if len(c) < 9:
                        continue
                    try:
                        sm.append(float(c[1])); mx.append(float(c[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm), reasons=sorted(reasons), source="nvidia-smi")
        return out


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------
def cpu_sample_frames(cores: int):
    """Bounded sample of the workload: per host core one 2-frame GOP (intra + inter) of a
    512x512 crop of the sequence, each core a different crop / GOP."""
    from mptc_b200.synth import make_frame
    crops = []
    for t in range(cores):
        g = t % (FRAMES // GOP)
        f0 = g * GOP
        x0 = (t * 192) % (W - 512) // 4 * 4
        y0 = (t * 128) % (H - 512) // 4 * 4
        pair = np.stack([make_frame(W, H, f0 + k)[y0:y0 + 512, x0:x0 + 512] for k in range(2)])
        crops.append(np.ascontiguousarray(pair))
    return crops


def run_cpu_sample(crops, kind: str, stages=None):
    """One GOP per thread (ThreadedCompressMultiUnique, codec.cpp:1781-1793, without its
    5-thread cap).  Returns seconds of wall time of stages A+B (DXT1 fit + index search).  With a
    dict in `stages` (reference kind only) the wavelet + arithmetic-coding stage C
    (EntropyEncode, codec.cpp:1115-1158) is then run and timed the same way, and the dict
    receives the wall time of C and the per-stage CPU seconds summed over the threads."""
    kept = []
    if kind == "reference":
        from oracle import ref

        def work(pair):
            prev = None
            for i in range(pair.shape[0]):
                fr = ref.RefFrame(pair[i], i == 0, SA, THR)   # DXTImage ctor (stb fit)
                fr.reencode(prev)                             # DXTImage::Reencode
                prev = fr
                if stages is not None:
                    kept.append(fr)
        ref.lib()  # load + silence before threading
        ref.RefFrame(crops[0][0][:16, :16], True, 1, THR)     # stb table init is not thread safe
    else:
        from oracle import port

        def work(pair):
            port.encode_gops(pair, GOP, SA, THR, 1, want_outputs=False)
        port.lib()
    threads = [threading.Thread(target=work, args=(c,)) for c in crops]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    t_ab = time.perf_counter() - t0
    if stages is not None and kept:
        per = (len(kept) + len(crops) - 1) // len(crops)
        chunks = [kept[i:i + per] for i in range(0, len(kept), per)]
        threads = [threading.Thread(target=lambda fs: [f.entropy_payload() for f in fs], args=(c,)) for c in chunks]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        stages["wall_c_s"] = time.perf_counter() - t0
        tt = [f.times() for f in kept]
        stages["cpu_s"] = {"fit": float(sum(t["fit_s"] for t in tt)), "search": float(sum(t["search_s"] for t in tt)),
                           "wavelet_entropy": float(sum(t["entropy_s"] for t in tt))}
    return t_ab


def cpu_kind():
    from oracle import ref
    return "reference" if ref.available() else "port"


def cpu_baseline(steps: int = 1):
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    crops = cpu_sample_frames(cores)
    pix = sum(c.shape[0] * c.shape[1] * c.shape[2] for c in crops)
    stages = {} if kind == "reference" else None
    times = [run_cpu_sample(crops, kind, stages if i == steps - 1 else None) for i in range(steps)]
    t = float(np.mean(times))
    out = {"value": pix / t / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{cores} threads x one 2-frame GOP (intra+inter) of a 512x512 crop of the {W}x{H} sequence, "
                     f"stages fit+search (DXTImage ctor + Reencode), sa={SA} thr={THR}; {t:.2f} s/step",
           "per_core": pix / t / 1e6 / cores}
    if stages and "cpu_s" in stages:   # stage A / B / C separately
        cs = stages["cpu_s"]
        out["stages_mpixel_per_s_per_core"] = {k: (pix / 1e6 / v if v > 0 else None) for k, v in cs.items()}
        out["with_wavelet_entropy"] = {"value": pix / (t + stages["wall_c_s"]) / 1e6, "unit": UNIT,
                                       "note": "stages A+B+C (EntropyEncode per frame, codec.cpp:1115-1158) on the same threads"}
    return out, times


def reference_arm(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    crops = cpu_sample_frames(cores)
    pix = sum(c.shape[0] * c.shape[1] * c.shape[2] for c in crops)
    for _ in range(min(args.warmup, 1)):   # CPU code has no warm-up effects worth minutes of wall time
        run_cpu_sample(crops, kind)
    times = [run_cpu_sample(crops, kind) for _ in range(args.steps)]
    t = float(np.mean(times))
    val = pix / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
        "config": base_config(max(args.gpus, 1)),
        "reference_sample": "reference CPU encoder; each step is a bounded sample of the workload named in config",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{cores} threads x one 2-frame GOP of a 512x512 crop, fit+search, sa={SA} thr={THR}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=claim_stdout(), flush=True)


def ncu_evidence(stage: str):
    """Counters of the dominant kernel from the committed ncu capture (profiles/ncu_counters.json,
    written by profiles/ncu_summary.py --json).  The capture is stamped with the hash of the kernel's
    sources; a capture of other sources says nothing about the kernel that was just timed and is
    reported as stale instead of being copied into the line."""
    from mptc_b200.build import kernel_source_sha
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    try:
        with open(p) as f:
            ev = json.load(f).get(stage)
    except Exception:
        return None, "profiles/ncu_counters.json missing"
    if not ev:
        return None, f"no capture of the {stage} kernel"
    have, want = ev.get("source_sha16"), kernel_source_sha(stage)
    if have != want:
        return None, f"capture is of other kernel sources (sha {have}, now {want}); re-run profiles/run_profile.sh"
    keep = ("kernel", "frames_per_launch", "source", "source_sha16", "gpu__time_duration.sum", "dram_bytes_per_frame",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    return {k: ev[k] for k in keep if k in ev}, None


def parity_record(out, first_frame: int, what: str, w: int, h: int, sa: int, thr: int):
    """Checks the outputs of the step that was just timed against the committed full-GOP fixture of the
    unmodified reference for this configuration (tests/golden/gen_golden_full.py).  Hashing only --
    no oracle code runs here."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import compare_with_full_fixture, full_fixture_name, load
    name = full_fixture_name(w, h, sa, thr, GOP)
    if name is None:
        return {"fixture": None, "frames": 0, "equal": None, "note": "no reference fixture for this configuration"}
    g = load(name)
    n, bad = compare_with_full_fixture(g, out["blocks"], out["motion"], out["unique"], out["n_unique"], first_frame=first_frame)
    rec = {"fixture": f"tests/golden/{name}.npz", "frames": n, "equal": (not bad) if n else None,
           "compared": "per-frame SHA-256 of final blocks, motion bytes and unique palette of " + what +
                       " vs the unmodified reference (oracle/_ref) on the same frames",
           "not_covered": "endpoint planes at 1080p/4K: the reference's wavelet is undefined unless the plane sizes are "
                          "multiples of 64 blocks (image_processing.h:293-294); planes are pinned on 256-multiple fixtures"}
    if bad:
        rec["mismatches"] = bad[:8]
    return rec


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def gen_frames(pin, w, h, first_frame, kind="clean", seed=1234):
    """Fills a pinned (n, h, w, 3) array with the synthetic sequence (numpy releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from mptc_b200.synth import make_frame

    def gen(f):
        fr = make_frame(w, h, first_frame + f, seed)
        if kind == "noisy":   # camera-like noise of +-24 on top of the sequence: ~distinct index words
            rng = np.random.default_rng(1000 + first_frame + f)
            fr = np.clip(fr.astype(np.int16) + rng.integers(-24, 25, size=fr.shape, dtype=np.int16), 0, 255).astype(np.uint8)
        pin[f] = fr
    with ThreadPoolExecutor(max(1, len(os.sched_getaffinity(0)))) as pool:
        list(pool.map(gen, range(pin.shape[0])))


def gpu_arm(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from mptc_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the encoder hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = capi.bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else {"numa_node": None}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cores = len(os.sched_getaffinity(0))
    host_threads = max(1, min(cores, (os.cpu_count() or 1) // max(world, 1)))   # the ranks of one box share its cores
    ctx = capi.Context(local_rank)
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    issue_peak = SM_COUNT * SCHEDULERS * LANES * sm_max_mhz * 1e6 / 1e12

    def measure(w, h, n_frames, first_frame, sa, thr, kind="clean", steps=args.steps, warmup=args.warmup,
                want_stream=True, want_parity=False, sampler=None):
        """One workload through the three timed paths.  Returns a dict of raw results (this rank)."""
        nb = (w // 4) * (h // 4)
        pbw, pbh = (w // 4 + 63) // 64 * 64, (h // 4 + 63) // 64 * 64
        pin_frames = capi.PinnedArray((n_frames, h, w, 3), np.uint8)
        gen_frames(pin_frames.array, w, h, first_frame, kind)
        frames = pin_frames.array
        pins = {"blocks": capi.PinnedArray((n_frames, nb), np.uint64), "motion": capi.PinnedArray((n_frames, 2 * nb), np.uint8),
                "unique": capi.PinnedArray((n_frames, nb), np.uint32), "n_unique": capi.PinnedArray((n_frames,), np.uint32),
                "planes": capi.PinnedArray((n_frames, 6, pbh, pbw), np.uint8)}
        out = {k: v.array for k, v in pins.items()}
        ctx.seq_reserve(w, h, n_frames)
        ctx.seq_upload(frames)
        ctx.sync()
        r = {"frames": frames, "out": out, "pins": (pin_frames, pins), "nb": nb, "pixels": n_frames * w * h}
        # ---- device-resident: inputs in HBM when the timed region starts -----------------------
        for _ in range(warmup):
            ctx.seq_encode(0, n_frames, sa, thr, GOP)
        ctx.sync()
        barrier()
        if sampler:
            sampler.start()   # nvidia-smi clocks / throttle reasons DURING the timed regions (resident + end to end)
        launches0 = ctx.launches
        dev_ms, stage_ms = [], {k: 0.0 for k in ("fit", "inter", "intra", "compact", "planes")}
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.seq_encode(0, n_frames, sa, thr, GOP)
            dev_ms.append(ctx.last_encode_ms("total"))   # synchronises on the step's end event
            for k in stage_ms:
                stage_ms[k] += ctx.last_encode_ms(k) / steps
        barrier()
        r["wall_ms"] = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        r["launches"] = ctx.launches - launches0
        r["step_ms"] = max_over_ranks(float(np.mean(dev_ms)))
        r["stage_ms"] = stage_ms
        r["work"] = ctx.last_work_count()
        # ---- end to end: host (pinned) frames in, host results out -------------------------------
        e2e = lambda: ctx.encode_sequence(frames, sa, thr, GOP, out=out)   # noqa: E731  (returns after the D2H copies)
        for _ in range(min(warmup, 3)):
            e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e()
        barrier()
        r["e2e_ms"] = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        if sampler:
            r["clocks"] = sampler.stop()
        r["h2d_bytes"] = int(frames.nbytes)
        # blocks, motion, planes are copied whole; of `unique` only the n_unique words of each frame travel
        r["d2h_bytes"] = int(out["blocks"].nbytes + out["motion"].nbytes + out["planes"].nbytes + out["n_unique"].nbytes +
                             4 * int(out["n_unique"].sum()))
        r["n_unique_sum"] = int(out["n_unique"].sum())
        if want_parity:
            r["parity"] = parity_record(out, first_frame, "the timed end-to-end step", w, h, sa, thr)
        # ---- end to end incl. the host arithmetic coder (stream bytes out), overlapped with the GPU ----
        if want_stream:
            stream_buf = np.empty(frames.nbytes // 4 + (1 << 20), dtype=np.uint8)   # the caller's output buffer, reused
            for _ in range(2):
                r["stream_bytes"] = len(capi.encode_stream(ctx, frames, sa, thr, GOP, host_threads, out=stream_buf)[0])
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                capi.encode_stream(ctx, frames, sa, thr, GOP, host_threads, out=stream_buf)
            barrier()
            r["stream_ms"] = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        return r

    def release(r):
        pf, pins = r.pop("pins")
        r.pop("frames", None); r.pop("out", None)
        pf.free()
        for v in pins.values():
            v.free()

    # ============================ the headline workload ==========================================
    m = measure(W, H, FRAMES, FRAMES * rank, SA, THR, want_parity=(rank == 0), sampler=ClockSampler(local_rank))
    clocks = m["clocks"]
    frames, out, nb = m["frames"], m["out"], m["nb"]

    # ---- decoder side: the stream just written back to DXT1 blocks ----------------
    stream = capi.encode_stream(ctx, frames, SA, THR, GOP, host_threads)[0]
    dec_pin = capi.PinnedArray((FRAMES, nb), np.uint64)
    ctx.seq_encode(0, FRAMES, SA, THR, GOP)           # leaves motion / unique / planes on the device
    ctx.sync()
    dec_ms = {k: 0.0 for k in ("total", "words", "planes", "rgb")}
    for it in range(2 + args.steps):
        ctx.seq_decode(0, FRAMES, SA, GOP, rgb=True)  # device-resident symbols -> blocks + RGB
        if it >= 2:
            for k in dec_ms:
                dec_ms[k] += ctx.last_decode_ms(k) / args.steps
    ctx.sync()
    dec_stats = None
    capi.decode_stream(ctx, stream, threads=host_threads, blocks_out=dec_pin.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dec_blocks, _, dec_stats = capi.decode_stream(ctx, stream, threads=host_threads, blocks_out=dec_pin.array)
    barrier()
    dec_stream_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    dec_ok = bool(np.array_equal(dec_blocks, out["blocks"]))
    for k in dec_ms:
        dec_ms[k] = max_over_ranks(dec_ms[k])

    # ---- kernel attribution pass: one GOP lane, so that kernels do not overlap and the CUDA events
    # around each launch measure that kernel alone (in the timed region above, kernels of different
    # lanes share the GPU and their event intervals include each other's time) ---------------------
    ctx.seq_upload(frames)
    ctx.set_schedule(1, 0, 0)
    for _ in range(2):
        ctx.seq_encode(0, FRAMES, SA, THR, GOP)
    ctx.sync()
    serial_ms, serial_total = {k: 0.0 for k in m["stage_ms"]}, 0.0
    for _ in range(args.steps):
        ctx.seq_encode(0, FRAMES, SA, THR, GOP)
        serial_total += ctx.last_encode_ms("total") / args.steps
        for k in serial_ms:
            serial_ms[k] += ctx.last_encode_ms(k) / args.steps
    ctx.set_schedule(0, 0, 0)
    dec_pin.free()
    release(m)

    # ============================ other content / other geometry ====================================
    legs = {}
    if not args.no_extra_legs and args.workload == "1080p60":
        def leg(w, h, n_frames, first, sa, thr, kind, stream, parity=False):
            x = measure(w, h, n_frames, first, sa, thr, kind, steps=max(2, args.steps // 2), warmup=3, want_stream=stream,
                        want_parity=parity and rank == 0)
            pix = x["pixels"] * world
            o = {"workload": f"{w}x{h} synthetic{' + noise(+-24)' if kind == 'noisy' else ''} x{n_frames} frames/GPU, "
                             f"search_area={sa}, err_threshold={thr}, gop={GOP}",
                 "value": pix / (x["step_ms"] * 1e-3) / 1e6, "ms_per_step": x["step_ms"],
                 "e2e": pix / (x["e2e_ms"] * 1e-3) / 1e6, "unit": UNIT, "n_unique": x["n_unique_sum"],
                 # distinct index words a target block's window holds on average = evaluations per target block
                 "evaluations_per_target_block": x["work"]["inter_evals"] / max(1, (h // 4) * (w // 4) * (n_frames - (n_frames + GOP - 1) // GOP))}
            if stream:
                o["e2e_stream"] = pix / (x["stream_ms"] * 1e-3) / 1e6
            if "parity" in x:
                o["parity"] = x["parity"]
            release(x)
            return o
        # word-diverse content: the de-duplicating kernels live on word re-use
        legs["robustness"] = {
            "noisy_thr50": leg(W, H, GOP, 0, SA, THR, "noisy", False),
            "clean_thr0": leg(W, H, GOP, 0, SA, 0, "clean", False),
            "note": "one GOP of 15 frames each; same kernels, content / threshold that keep the frames' own ~distinct index words",
        }
        # BASELINE configs[2] geometry: 4K, host entropy coding overlapped (4 GOPs per GPU)
        legs["4k"] = leg(3840, 2160, 4 * GOP, 4 * GOP * rank, SA, THR, "clean", True, parity=True)
        # BASELINE configs[4]: 4K x 600 frames in all, GOP-sharded over the GPUs (strong scaling).  Only under
        # torchrun (the whole job on one GPU is 15 GB of frames); the single-GPU yardstick is the `4k` leg
        # of the N = 1 run (same geometry, >= 4 GOPs per GPU, no collective).
        n_gops_total = 600 // GOP
        if world > 1 and n_gops_total % world == 0:
            per_rank = n_gops_total // world * GOP
            x = measure(3840, 2160, per_rank, per_rank * rank, SA, THR, "clean", steps=2, warmup=1, want_stream=True,
                        want_parity=(rank == 0))
            pix = 600 * 3840 * 2160
            legs["4k600_strong"] = {
                "workload": f"3840x2160 synthetic x600 frames in all = {per_rank} frames/GPU, search_area={SA}, "
                            f"err_threshold={THR}, gop={GOP}", "scaling": "strong", "unit": UNIT,
                "value": pix / (x["step_ms"] * 1e-3) / 1e6, "ms_per_step": x["step_ms"],
                "e2e": pix / (x["e2e_ms"] * 1e-3) / 1e6, "e2e_ms_per_step": x["e2e_ms"],
                "e2e_stream": pix / (x["stream_ms"] * 1e-3) / 1e6,
                "h2d_gb_per_s_per_rank": x["h2d_bytes"] / (x["e2e_ms"] * 1e-3) / 1e9,
                "d2h_gb_per_s_per_rank": x["d2h_bytes"] / (x["e2e_ms"] * 1e-3) / 1e9,
                "parity": x.get("parity"),
                "single_gpu_reference": "legs.4k of the N=1 run (Mpixel/s of one GPU on the same geometry)"}
            release(x)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pixels_total = m["pixels"] * world
    step_ms, e2e_ms, stream_ms, stage_ms = m["step_ms"], m["e2e_ms"], m["stream_ms"], m["stage_ms"]
    dominant = max(("inter", "intra"), key=lambda k: serial_ms[k])
    work = m["work"]
    n_gops = FRAMES // GOP
    k_launches = (GOP - 1) if dominant == "inter" else GOP   # one launch covers frame k of every GOP
    k_ms = serial_ms[dominant]
    evals, scanned = work[f"{dominant}_evals"], work[f"{dominant}_scanned"]
    nominal = work[f"nominal_{dominant}"]
    slots = evals * ALGO_SLOTS_PER_CANDIDATE + scanned * SCAN_SLOTS_PER_POSITION
    achieved = slots / (k_ms * 1e-3) / 1e12
    # algorithmic HBM bytes of the search kernels: 72 B/block in + 20 B per
    # distinct candidate + ~14 B/block out
    frames_k = n_gops * ((GOP - 1) if dominant == "inter" else GOP)
    algo_bytes = frames_k * nb * (72 + 20 + 14)
    ev, stale = ncu_evidence(dominant)
    traffic = ev["dram_bytes_per_frame"] * n_gops if ev and "dram_bytes_per_frame" in ev else None
    roofline = {
        "kernel": ("k_inter_search_wide<int8_t>" if (SA <= 16 and THR < 127) else "k_inter_search_tiled") if dominant == "inter" else "k_intra_rows",
        "measured": "CUDA events around each launch in a one-lane pass of the same workload (kernels serialised)",
        "share_of_step": k_ms / serial_total, "serial_step_ms": serial_total, "serial_stage_ms": serial_ms,
        "bound": "issue", "achieved": achieved, "peak": issue_peak, "unit": "Tslot/s", "frac": achieved / issue_peak,
        "peak_source": f"148 SMs x 4 schedulers x 32 lanes x {sm_max_mhz:.0f} MHz (clocks.max.sm, {peak_src}); issue-slot roofline per SURVEY.md 8(d)",
        "work": "EXECUTED work, counted by the kernel (one atomicAdd per tile): (index word, target block) evaluations x "
                f"{ALGO_SLOTS_PER_CANDIDATE} slots (SURVEY.md 8d) + window positions scanned x {SCAN_SLOTS_PER_POSITION} slots",
        "evaluations_per_step": evals, "positions_scanned_per_step": scanned,
        "units_per_launch": evals / max(k_launches, 1), "launches_per_step": k_launches,
        "avg_launch_ms": k_ms / max(k_launches, 1), "algo_slots_per_unit": ALGO_SLOTS_PER_CANDIDATE,
        "scan_slots_per_position": SCAN_SLOTS_PER_POSITION,
        "evaluations_per_target_block": work["inter_evals"] / max(1, n_gops * (GOP - 1) * nb),
        # what the same time would score if every window position were evaluated, as the reference does
        # and as SURVEY.md 8(d) counts: the de-duplication's gain, not a roofline fraction
        "nominal_candidates_per_step": nominal,
        "dedup_speedup": nominal * ALGO_SLOTS_PER_CANDIDATE / max(slots, 1),
        "traffic": traffic,
        "hbm": {"achieved": algo_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": algo_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src},
        "overlapped_stage_ms_per_step": stage_ms,
        "candidates_per_step": {"inter": work["nominal_inter"], "intra": work["nominal_intra"]},
        "ncu": ev,
    }
    if stale:
        roofline["ncu_stale"] = stale
    if ev and "smsp__issue_active.avg.pct_of_peak_sustained_active" in ev:
        roofline["issue_active_ncu"] = ev["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100.0
    if clocks.get("sm_mhz"):
        roofline["frac_at_sampled_clock"] = achieved / (SM_COUNT * SCHEDULERS * LANES * clocks["sm_mhz"] * 1e6 / 1e12)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, cpu_times = cpu_baseline(3)
        cpu["steps"] = len(cpu_times)
        cpu["step_s"] = [round(t, 3) for t in cpu_times]

    cfg = base_config(world)
    line = {
        "metric": METRIC, "value": pixels_total / (step_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": SCALING, "vs_baseline": None, "dtype": "int32+f32", "data": "synthetic",
        "config": cfg,
        "wall_ms_per_step": m["wall_ms"],
        "e2e": {"value": pixels_total / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                "h2d_gb_per_s_per_rank": m["h2d_bytes"] / (e2e_ms * 1e-3) / 1e9,
                "d2h_gb_per_s_per_rank": m["d2h_bytes"] / (e2e_ms * 1e-3) / 1e9, "numa": numa},
        "e2e_stream": {"value": pixels_total / (stream_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": stream_ms,
                       "host_threads": host_threads, "stream_bytes": m["stream_bytes"],
                       "note": "e2e + the host arithmetic coder and stream assembly (mptc_encode_stream), coder overlapped with the GPU"},
        "parity": m.get("parity"),
        "gpu_launches": int(m["launches"]), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "checksum_n_unique": m["n_unique_sum"],
        "legs": legs,
        "decode": {
            "note": "decoder side (SURVEY.md 8f-2), same frames: device = symbols resident in HBM -> DXT1 blocks + RGB "
                    "pictures (CUDA events); stream = mptc_decode_stream from the stream bytes to host DXT1 blocks, host "
                    "arithmetic decoder included",
            "device": {"value": pixels_total / (dec_ms["total"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dec_ms["total"],
                       "stage_ms": dec_ms},
            "rgb_kernel_hbm": {"kernel": "k_dxt1_to_rgb", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak,
                               "achieved": FRAMES * nb * 56 / (dec_ms["rgb"] * 1e-3) / 1e9,
                               "frac": FRAMES * nb * 56 / (dec_ms["rgb"] * 1e-3) / 1e9 / hbm_peak,
                               "algo_bytes_per_block": 56},
            "stream": {"value": pixels_total / (dec_stream_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dec_stream_ms,
                       "host_threads": host_threads, "entropy_ms": dec_stats.entropy_ms if dec_stats else None,
                       "symbols": int(dec_stats.symbols) if dec_stats else None},
            "round_trip_equal": dec_ok,
        },
    }
    print(json.dumps(line), file=claim_stdout(), flush=True)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything else a library prints there (NCCL's version
    banner, for one) is sent to stderr by pointing fd 1 at fd 2 and keeping the real stdout aside."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mptc_b200", choices=["mptc_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the robustness and 4K legs of the default line")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the ranks to their GPU's NUMA node (N > 1)")
    ap.add_argument("--workload", default="1080p60", choices=sorted(WORKLOADS),
                    help="1080p60 = BASELINE.json configs[1] (default, the headline); 4k120 = configs[2]; "
                         "4k600 = configs[4] (600 frames in all, GOP-sharded over the GPUs: strong scaling)")
    ap.add_argument("--search-area", type=int, default=SA)
    ap.add_argument("--err-threshold", type=int, default=THR)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    set_workload(args.workload, args.search_area, args.err_threshold, world)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
