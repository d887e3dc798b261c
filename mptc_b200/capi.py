"""ctypes binding of libmptc_b200.so -- the C-ABI declared in include/mptc_gpu.h.

There is no CPU fallback: loading fails loudly when the CUDA library has not been built,
and every compute call fails loudly when no GPU is present."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MPTC_OK = 0
STAGES = {"total": 0, "fit": 1, "inter": 2, "intra": 3, "compact": 4, "planes": 5}

# every symbol include/mptc_gpu.h declares (tests check the library exports all of them)
EXPORTS = [
    "mptc_gpu_create", "mptc_gpu_destroy", "mptc_gpu_last_error", "mptc_gpu_launch_count",
    "mptc_gpu_dxt1_fit", "mptc_gpu_reencode", "mptc_gpu_endpoint_planes",
    "mptc_gpu_seq_reserve", "mptc_gpu_seq_upload", "mptc_gpu_seq_encode", "mptc_gpu_seq_download",
    "mptc_gpu_sync", "mptc_gpu_last_encode_ms", "mptc_gpu_encode_sequence",
    "mptc_gpu_host_alloc", "mptc_gpu_host_free", "mptc_gpu_last_candidate_count", "mptc_gpu_set_schedule",
    "mptc_gpu_encode_sequence_async", "mptc_gpu_wait_frame", "mptc_gpu_wait",
    "mptc_gpu_decode_sequence", "mptc_gpu_seq_decode_upload", "mptc_gpu_seq_decode",
    "mptc_gpu_seq_decode_download", "mptc_gpu_last_decode_ms", "mptc_gpu_last_work_count", "mptc_gpu_inter_pixel_search",
]
DECODE_STAGES = {"total": 0, "words": 1, "planes": 2, "rgb": 3}


class Params(C.Structure):
    _fields_ = [("search_area", C.c_int), ("err_threshold", C.c_int), ("gop", C.c_int)]


class MptcError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Loads libmptc_b200.so; raises if it is missing (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        raise MptcError(f"{_build.LIB} is missing: run `python -m mptc_b200.build` (needs nvcc). "
                        "There is no CPU fallback for the encoder hot path.")
    L = C.CDLL(_build.LIB)
    vp, ci = C.c_void_p, C.c_int
    L.mptc_gpu_create.argtypes = [ci, C.POINTER(vp)]
    L.mptc_gpu_destroy.argtypes = [vp]
    L.mptc_gpu_last_error.restype = C.c_char_p
    L.mptc_gpu_last_error.argtypes = [vp]
    L.mptc_gpu_launch_count.restype = C.c_uint64
    L.mptc_gpu_launch_count.argtypes = [vp]
    L.mptc_gpu_dxt1_fit.argtypes = [vp, vp, ci, ci, vp]
    L.mptc_gpu_reencode.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp]
    L.mptc_gpu_endpoint_planes.argtypes = [vp, vp, ci, ci, vp]
    L.mptc_gpu_seq_reserve.argtypes = [vp, ci, ci, ci]
    L.mptc_gpu_seq_upload.argtypes = [vp, vp, ci, ci]
    L.mptc_gpu_seq_encode.argtypes = [vp, ci, ci, C.POINTER(Params)]
    L.mptc_gpu_seq_download.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp]
    L.mptc_gpu_sync.argtypes = [vp]
    L.mptc_gpu_last_encode_ms.argtypes = [vp, ci, C.POINTER(C.c_float)]
    L.mptc_gpu_encode_sequence.argtypes = [vp, vp, ci, ci, ci, C.POINTER(Params), vp, vp, vp, vp, vp]
    L.mptc_gpu_encode_sequence_async.argtypes = [vp, vp, ci, ci, ci, C.POINTER(Params), vp, vp, vp, vp, vp]
    L.mptc_gpu_wait_frame.argtypes = [vp, ci]
    L.mptc_gpu_wait.argtypes = [vp]
    L.mptc_gpu_host_alloc.restype = vp
    L.mptc_gpu_host_alloc.argtypes = [C.c_size_t]
    L.mptc_gpu_host_free.argtypes = [vp]
    L.mptc_gpu_last_candidate_count.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mptc_gpu_set_schedule.argtypes = [vp, ci, ci, ci]
    L.mptc_gpu_last_work_count.argtypes = [vp, C.POINTER(C.c_uint64), ci]
    L.mptc_gpu_inter_pixel_search.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp]
    L.mptc_gpu_decode_sequence.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, ci, ci, ci, ci, ci, vp, vp]
    L.mptc_gpu_seq_decode_upload.argtypes = [vp, ci, ci, vp, vp, vp, C.c_size_t, vp]
    L.mptc_gpu_seq_decode.argtypes = [vp, ci, ci, ci, ci, ci]
    L.mptc_gpu_seq_decode_download.argtypes = [vp, ci, ci, vp, vp]
    L.mptc_gpu_last_decode_ms.argtypes = [vp, ci, C.POINTER(C.c_float)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data


def bind_to_gpu_numa_node(device: int) -> dict:
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the
    page-locked frame / result buffers it allocates next and its arithmetic-coder threads are local
    to the GPU's PCIe root: with eight ranks on a two-socket box, half of the H2D/D2H traffic otherwise
    crosses the socket interconnect.  Returns what was done (for the bench line); never fails."""
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception as e:   # containers without sysfs, single-node boxes, ...
        info["error"] = str(e)[:80]
    return info


class PinnedArray:
    """numpy view over page-locked host memory from mptc_gpu_host_alloc."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = load().mptc_gpu_host_alloc(max(self.nbytes, 1))
        if not self._p:
            raise MptcError("mptc_gpu_host_alloc failed")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            load().mptc_gpu_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One mptc_gpu_ctx: one per (host thread, GPU)."""

    def __init__(self, device: int = 0):
        self._L = load()
        p = C.c_void_p()
        r = self._L.mptc_gpu_create(device, C.byref(p))
        if r != MPTC_OK:
            raise MptcError(f"mptc_gpu_create(device={device}) failed with {r}: no usable CUDA device "
                            "(the encoder hot path has no CPU fallback)")
        self._p = p
        self.device = device
        self.w = self.h = self.nb = 0

    def close(self):
        if getattr(self, "_p", None):
            self._L.mptc_gpu_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, r):
        if r != MPTC_OK:
            raise MptcError(f"mptc_gpu error {r}: {self._L.mptc_gpu_last_error(self._p).decode()}")

    @property
    def launches(self) -> int:
        return int(self._L.mptc_gpu_launch_count(self._p))

    # ---- single frame ------------------------------------------------------------------
    def dxt1_fit(self, rgb: np.ndarray) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w = rgb.shape[:2]
        out = np.empty((h // 4) * (w // 4), dtype=np.uint64)
        self._check(self._L.mptc_gpu_dxt1_fit(self._p, rgb.ctypes.data, w, h, out.ctypes.data))
        return out

    def reencode(self, rgb, is_intra, search_area, err_threshold, prev_blocks=None):
        """-> dict(initial, blocks, motion, unique)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w = rgb.shape[:2]
        nb = (h // 4) * (w // 4)
        initial = np.empty(nb, dtype=np.uint64)
        blocks = np.empty(nb, dtype=np.uint64)
        motion = np.empty(2 * nb, dtype=np.uint8)
        unique = np.empty(nb, dtype=np.uint32)
        nu = C.c_uint32(0)
        if prev_blocks is not None:
            prev_blocks = np.ascontiguousarray(prev_blocks, dtype=np.uint64)
        self._check(self._L.mptc_gpu_reencode(self._p, rgb.ctypes.data, w, h, int(is_intra), search_area,
                                              err_threshold, _ptr(prev_blocks), initial.ctypes.data,
                                              blocks.ctypes.data, motion.ctypes.data, unique.ctypes.data,
                                              C.addressof(nu)))
        return {"initial": initial, "blocks": blocks, "motion": motion, "unique": unique[: nu.value].copy()}

    def inter_pixel_search(self, rgb, search_area, prev_blocks, cur_blocks=None):
        """DXTImage::InterPixelSearch for every block -> dict(min_err, motion, index, reassigned)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w = rgb.shape[:2]
        nb = (h // 4) * (w // 4)
        prev_blocks = np.ascontiguousarray(prev_blocks, dtype=np.uint64)
        if cur_blocks is not None:
            cur_blocks = np.ascontiguousarray(cur_blocks, dtype=np.uint64)
        out = {"min_err": np.empty(nb, np.int32), "motion": np.empty(2 * nb, np.uint8), "index": np.empty(nb, np.uint32),
               "reassigned": np.empty(nb, np.uint8)}
        self._check(self._L.mptc_gpu_inter_pixel_search(self._p, rgb.ctypes.data, w, h, search_area, _ptr(cur_blocks),
                                                        prev_blocks.ctypes.data, out["min_err"].ctypes.data,
                                                        out["motion"].ctypes.data, out["index"].ctypes.data,
                                                        out["reassigned"].ctypes.data))
        self.w = self.h = self.nb = 0   # the reserved sequence was re-purposed
        return out

    def endpoint_planes(self, blocks, bw, bh) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, dtype=np.uint64)
        pbw, pbh = (bw + 63) // 64 * 64, (bh + 63) // 64 * 64
        out = np.empty((6, pbh, pbw), dtype=np.uint8)
        self._check(self._L.mptc_gpu_endpoint_planes(self._p, blocks.ctypes.data, bw, bh, out.ctypes.data))
        return out

    # ---- sequences -----------------------------------------------------------------------
    def seq_reserve(self, w, h, n_frames):
        self._check(self._L.mptc_gpu_seq_reserve(self._p, w, h, n_frames))
        self.w, self.h, self.nb = w, h, (w // 4) * (h // 4)
        self.pbw, self.pbh = (w // 4 + 63) // 64 * 64, (h // 4 + 63) // 64 * 64

    def seq_upload(self, frames: np.ndarray, first=0):
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        self._check(self._L.mptc_gpu_seq_upload(self._p, frames.ctypes.data, first, frames.shape[0]))

    def seq_encode(self, first, count, search_area, err_threshold, gop):
        p = Params(search_area, err_threshold, gop)
        self._check(self._L.mptc_gpu_seq_encode(self._p, first, count, C.byref(p)))

    def seq_download(self, first, count, want=("blocks", "initial", "motion", "unique", "planes"), into=None):
        nb = self.nb
        out = dict(into) if into else {}
        if "blocks" in want and "blocks" not in out:
            out["blocks"] = np.empty((count, nb), dtype=np.uint64)
        if "initial" in want and "initial" not in out:
            out["initial"] = np.empty((count, nb), dtype=np.uint64)
        if "motion" in want and "motion" not in out:
            out["motion"] = np.empty((count, 2 * nb), dtype=np.uint8)
        if "unique" in want and "unique" not in out:
            out["unique"] = np.empty((count, nb), dtype=np.uint32)
            out["n_unique"] = np.empty(count, dtype=np.uint32)
        if "planes" in want and "planes" not in out:
            out["planes"] = np.empty((count, 6, self.pbh, self.pbw), dtype=np.uint8)
        self._check(self._L.mptc_gpu_seq_download(self._p, first, count, _ptr(out.get("blocks")),
                                                  _ptr(out.get("initial")), _ptr(out.get("motion")),
                                                  _ptr(out.get("unique")), _ptr(out.get("n_unique")),
                                                  _ptr(out.get("planes"))))
        return out

    def sync(self):
        self._check(self._L.mptc_gpu_sync(self._p))

    def set_schedule(self, lanes=0, wave_rows_intra=0, wave_rows_inter=0):
        """GOP lanes (streams) per encode call and CTAs per frame of the intra wavefront; 0 = automatic."""
        self._check(self._L.mptc_gpu_set_schedule(self._p, lanes, wave_rows_intra, wave_rows_inter))

    def last_encode_ms(self, stage="total") -> float:
        ms = C.c_float(0)
        self._check(self._L.mptc_gpu_last_encode_ms(self._p, STAGES[stage], C.byref(ms)))
        return float(ms.value)

    def last_candidate_count(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self._L.mptc_gpu_last_candidate_count(self._p, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def last_work_count(self) -> dict:
        """Work the search kernels executed in the last encode call (mptc_gpu_last_work_count)."""
        a = (C.c_uint64 * 8)()
        self._check(self._L.mptc_gpu_last_work_count(self._p, a, 8))
        keys = ("nominal_inter", "nominal_intra", "inter_evals", "inter_scanned", "intra_evals", "intra_scanned",
                "inter_tiles", "intra_groups")
        return {k: int(x) for k, x in zip(keys, a)}

    def wait_frame(self, frame: int):
        """Blocks until the results of `frame` of the last (async) encode are in the host buffers."""
        self._check(self._L.mptc_gpu_wait_frame(self._p, frame))

    def wait(self):
        self._check(self._L.mptc_gpu_wait(self._p))

    def encode_sequence(self, frames: np.ndarray, search_area, err_threshold, gop, out=None, planes=True,
                        wait=True):
        """End to end from host frames to host results (H2D + kernels + D2H).  wait=False returns as
        soon as the work is enqueued (then use wait_frame / wait; keep `frames` and `out` alive)."""
        assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
        n, h, w = frames.shape[:3]
        nb = (h // 4) * (w // 4)
        pbw, pbh = (w // 4 + 63) // 64 * 64, (h // 4 + 63) // 64 * 64
        if out is None:
            out = {"blocks": np.empty((n, nb), dtype=np.uint64), "motion": np.empty((n, 2 * nb), dtype=np.uint8),
                   "unique": np.empty((n, nb), dtype=np.uint32), "n_unique": np.empty(n, dtype=np.uint32)}
            if planes:
                out["planes"] = np.empty((n, 6, pbh, pbw), dtype=np.uint8)
        p = Params(search_area, err_threshold, gop)
        fn = self._L.mptc_gpu_encode_sequence if wait else self._L.mptc_gpu_encode_sequence_async
        self._check(fn(self._p, frames.ctypes.data, n, w, h, C.byref(p), _ptr(out.get("blocks")), _ptr(out["motion"]),
                       _ptr(out["unique"]), _ptr(out["n_unique"]), _ptr(out.get("planes"))))
        self.w, self.h, self.nb, self.pbw, self.pbh = w, h, nb, pbw, pbh
        return out

    # ---- decoder side ----------------------------------------------------------------------
    def decode_sequence(self, motion, unique, n_unique, planes, w, h, search_area, gop, packed=False, rgb=False):
        """Symbols -> DXT1 blocks [n, nb] (and RGB frames [n, h, w, 3] with rgb=True).  `unique` is
        [n, nb] in the encoder's layout, or with packed=True the unique words of all frames back to
        back."""
        motion = np.ascontiguousarray(motion, dtype=np.uint8)
        unique = np.ascontiguousarray(unique, dtype=np.uint32)
        n_unique = np.ascontiguousarray(n_unique, dtype=np.uint32)
        planes = np.ascontiguousarray(planes, dtype=np.uint8)
        n = n_unique.size
        nb = (h // 4) * (w // 4)
        blocks = np.empty((n, nb), dtype=np.uint64)
        pix = np.empty((n, h, w, 3), dtype=np.uint8) if rgb else None
        self._check(self._L.mptc_gpu_decode_sequence(self._p, motion.ctypes.data, unique.ctypes.data,
                                                     n_unique.ctypes.data, 0 if packed else nb, planes.ctypes.data,
                                                     n, w, h, search_area, gop, blocks.ctypes.data, _ptr(pix)))
        self.w, self.h, self.nb = w, h, nb
        self.pbw, self.pbh = (w // 4 + 63) // 64 * 64, (h // 4 + 63) // 64 * 64
        return (blocks, pix) if rgb else blocks

    def seq_decode(self, first, count, search_area, gop, rgb=False):
        """Decodes what is on the device (after seq_encode: the device-resident round trip)."""
        self._check(self._L.mptc_gpu_seq_decode(self._p, first, count, search_area, gop, int(rgb)))

    def seq_decode_download(self, first, count, rgb=False):
        blocks = np.empty((count, self.nb), dtype=np.uint64)
        pix = np.empty((count, self.h, self.w, 3), dtype=np.uint8) if rgb else None
        self._check(self._L.mptc_gpu_seq_decode_download(self._p, first, count, blocks.ctypes.data, _ptr(pix)))
        return (blocks, pix) if rgb else blocks

    def last_decode_ms(self, stage="total") -> float:
        ms = C.c_float(0)
        self._check(self._L.mptc_gpu_last_decode_ms(self._p, DECODE_STAGES[stage], C.byref(ms)))
        return float(ms.value)


# ---- host codec (include/mptc_codec.h) -----------------------------------------------------------
CODEC_EXPORTS = ["mptc_arith_encode", "mptc_frame_payload", "mptc_encode_stream", "mptc_assemble_stream",
                 "mptc_arith_decode", "mptc_arith_decode_multi", "mptc_stream_info", "mptc_decode_stream"]


class StreamHeader(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("gop", C.c_int), ("search_area", C.c_int),
                ("n_groups", C.c_int), ("n_frames", C.c_int), ("max_unique_bytes", C.c_uint32),
                ("max_comp_palette", C.c_uint32), ("max_comp_motion", C.c_uint32), ("max_comp_ep_y", C.c_uint32),
                ("max_comp_ep_c", C.c_uint32)]


class DecodeStats(C.Structure):
    _fields_ = [("header", StreamHeader), ("entropy_ms", C.c_double), ("total_ms", C.c_double),
                ("symbols", C.c_uint64)]


class StreamStats(C.Structure):
    _fields_ = [("max_unique_bytes", C.c_uint32), ("max_comp_palette", C.c_uint32), ("max_comp_motion", C.c_uint32),
                ("max_comp_ep_y", C.c_uint32), ("max_comp_ep_c", C.c_uint32), ("n_groups", C.c_uint32),
                ("gpu_ms", C.c_double), ("entropy_ms", C.c_double), ("total_ms", C.c_double),
                ("assemble_ms", C.c_double)]


def _codec():
    L = load()
    if not getattr(L, "_codec_ready", False):
        vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
        L.mptc_arith_encode.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
        L.mptc_frame_payload.argtypes = [vp, sz, vp, sz, C.c_uint32, ci, vp, sz, C.POINTER(sz), vp]
        L.mptc_encode_stream.argtypes = [vp, vp, ci, ci, ci, C.POINTER(Params), ci, vp, sz, C.POINTER(sz),
                                         C.POINTER(StreamStats)]
        L.mptc_assemble_stream.argtypes = [ci, ci, ci, C.POINTER(Params), vp, vp, vp, vp, ci, vp, sz, C.POINTER(sz),
                                           C.POINTER(StreamStats)]
        L.mptc_arith_decode.argtypes = [vp, sz, vp, sz]
        L.mptc_arith_decode_multi.argtypes = [C.c_int, vp, vp, vp, vp]
        L.mptc_stream_info.argtypes = [vp, sz, C.POINTER(StreamHeader)]
        L.mptc_decode_stream.argtypes = [vp, vp, sz, ci, vp, vp, C.POINTER(DecodeStats)]
        L._codec_ready = True
    return L


def arith_decode(code: bytes, n: int) -> np.ndarray:
    """EntropyDecode (codec.cpp:560-577): n byte symbols from an arithmetic-coded stream."""
    buf = np.frombuffer(code, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    r = _codec().mptc_arith_decode(buf.ctypes.data if buf.size else None, buf.size, out.ctypes.data, n)
    if r != MPTC_OK:
        raise MptcError(f"mptc_arith_decode failed: {r}")
    return out


def arith_decode_multi(codes, ns):
    """Up to eight independent streams decoded in one interleaved loop (mptc_arith_decode_multi)."""
    k = len(codes)
    bufs = [np.frombuffer(c, dtype=np.uint8) for c in codes]
    outs = [np.empty(n, dtype=np.uint8) for n in ns]
    code_p = (C.c_void_p * k)(*[b.ctypes.data if b.size else None for b in bufs])
    sym_p = (C.c_void_p * k)(*[o.ctypes.data for o in outs])
    nbytes = (C.c_size_t * k)(*[b.size for b in bufs])
    nn = (C.c_size_t * k)(*ns)
    r = _codec().mptc_arith_decode_multi(k, code_p, nbytes, sym_p, nn)
    if r != MPTC_OK:
        raise MptcError(f"mptc_arith_decode_multi failed: {r}")
    return outs


def stream_info(stream: bytes) -> StreamHeader:
    buf = np.frombuffer(stream, dtype=np.uint8)
    hdr = StreamHeader()
    r = _codec().mptc_stream_info(buf.ctypes.data, buf.size, C.byref(hdr))
    if r != MPTC_OK:
        raise MptcError(f"mptc_stream_info failed: {r}")
    return hdr


def decode_stream(ctx: "Context", stream, threads=1, rgb=False, blocks_out=None):
    """Stream bytes -> (blocks [n, nb], rgb [n, h, w, 3] or None, DecodeStats).  blocks_out: a
    caller-owned (ideally page-locked) uint64 [n, nb] array to decode into."""
    buf = np.frombuffer(stream, dtype=np.uint8) if not isinstance(stream, np.ndarray) else stream
    hdr = stream_info(buf)
    n, w, h = hdr.n_frames, hdr.width, hdr.height
    nb = (w // 4) * (h // 4)
    blocks = np.empty((n, nb), dtype=np.uint64) if blocks_out is None else blocks_out
    assert blocks.shape == (n, nb) and blocks.dtype == np.uint64 and blocks.flags["C_CONTIGUOUS"]
    pix = np.empty((n, h, w, 3), dtype=np.uint8) if rgb else None
    st = DecodeStats()
    r = _codec().mptc_decode_stream(ctx._p, buf.ctypes.data, buf.size, threads, blocks.ctypes.data, _ptr(pix),
                                    C.byref(st))
    if r != MPTC_OK:
        raise MptcError(f"mptc_decode_stream failed: {r}: {ctx._L.mptc_gpu_last_error(ctx._p).decode()}")
    ctx.w, ctx.h, ctx.nb = w, h, nb
    return blocks, pix, st


def arith_encode(sym: np.ndarray) -> bytes:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    cap = 2 * sym.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    r = _codec().mptc_arith_encode(sym.ctypes.data, sym.size, out.ctypes.data, cap, C.byref(n))
    if r != MPTC_OK:
        raise MptcError(f"mptc_arith_encode failed: {r}")
    return out[: n.value].tobytes()


def frame_payload(motion: np.ndarray, planes: np.ndarray, n_unique: int, threads: int = 1):
    """-> (payload bytes, sizes[5])."""
    motion = np.ascontiguousarray(motion, dtype=np.uint8)
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    nb = motion.size // 2
    ps = planes.size // 6
    cap = 2 * (motion.size + planes.size) + 1024
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    sizes = np.zeros(5, dtype=np.uint32)
    r = _codec().mptc_frame_payload(motion.ctypes.data, nb, planes.ctypes.data, ps, n_unique, threads,
                                    out.ctypes.data, cap, C.byref(n), sizes.ctypes.data)
    if r != MPTC_OK:
        raise MptcError(f"mptc_frame_payload failed: {r}")
    return out[: n.value].tobytes(), sizes


def assemble_stream(w, h, search_area, err_threshold, gop, motion, unique, n_unique, planes, threads=1):
    """Host-only stream assembly from per-frame results -> (bytes, StreamStats)."""
    motion = np.ascontiguousarray(motion, dtype=np.uint8)
    unique = np.ascontiguousarray(unique, dtype=np.uint32)
    n_unique = np.ascontiguousarray(n_unique, dtype=np.uint32)
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    n = n_unique.size
    cap = 2 * (motion.size + planes.size) + 4 * unique.size + 4096
    out = np.empty(cap, dtype=np.uint8)
    nbytes = C.c_size_t(0)
    st = StreamStats()
    p = Params(search_area, err_threshold, gop)
    r = _codec().mptc_assemble_stream(n, w, h, C.byref(p), motion.ctypes.data, unique.ctypes.data,
                                      n_unique.ctypes.data, planes.ctypes.data, threads, out.ctypes.data, cap,
                                      C.byref(nbytes), C.byref(st))
    if r != MPTC_OK:
        raise MptcError(f"mptc_assemble_stream failed: {r}")
    return out[: nbytes.value].tobytes(), st


def encode_stream(ctx: "Context", frames: np.ndarray, search_area, err_threshold, gop, threads=1, out=None):
    """GPU hot path + host arithmetic coding -> (stream bytes, StreamStats).  With a caller-owned
    uint8 buffer in `out` (reused across calls) the stream is returned as a view into it."""
    assert frames.dtype == np.uint8 and frames.flags["C_CONTIGUOUS"]
    n, h, w = frames.shape[:3]
    own = out is None
    if own:
        out = np.empty(frames.nbytes // 2 + (1 << 20), dtype=np.uint8)
    cap = out.size
    nbytes = C.c_size_t(0)
    st = StreamStats()
    p = Params(search_area, err_threshold, gop)
    r = _codec().mptc_encode_stream(ctx._p, frames.ctypes.data, n, w, h, C.byref(p), threads, out.ctypes.data, cap,
                                    C.byref(nbytes), C.byref(st))
    if r != MPTC_OK:
        raise MptcError(f"mptc_encode_stream failed: {r}: {ctx._L.mptc_gpu_last_error(ctx._p).decode()}")
    return (out[: nbytes.value].tobytes() if own else out[: nbytes.value]), st
