"""ctypes binding of oracle/_ref/libmptc_ref.so (the UNMODIFIED reference, see
oracle/ref_wrap.cpp).  TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmptc_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.mptc_ref_frame_new.restype = C.c_void_p
        L.mptc_ref_frame_new.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.mptc_ref_frame_free.argtypes = [C.c_void_p]
        L.mptc_ref_num_blocks.argtypes = [C.c_void_p]
        L.mptc_ref_get_blocks.argtypes = [C.c_void_p, C.c_void_p]
        L.mptc_ref_reencode.argtypes = [C.c_void_p, C.c_void_p]
        L.mptc_ref_get_motion.argtypes = [C.c_void_p, C.c_void_p]
        L.mptc_ref_num_unique.argtypes = [C.c_void_p]
        L.mptc_ref_get_unique.argtypes = [C.c_void_p, C.c_void_p]
        L.mptc_ref_psnr_logical.restype = C.c_double
        L.mptc_ref_psnr_logical.argtypes = [C.c_void_p]
        L.mptc_ref_psnr_physical.restype = C.c_double
        L.mptc_ref_psnr_physical.argtypes = [C.c_void_p]
        L.mptc_ref_entropy_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mptc_ref_payload_planes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mptc_ref_get_times.argtypes = [C.c_void_p, C.c_void_p]
        L.mptc_ref_arith_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.mptc_ref_compress_multi_unique.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_int, C.c_uint, C.c_uint]
        L.mptc_ref_selfcheck_png.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.mptc_ref_write_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.mptc_ref_inter_pixel_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mptc_ref_inter_pixel_search_defined.argtypes = L.mptc_ref_inter_pixel_search.argtypes
        L.mptc_ref_quiet()
        _lib = L
    return _lib


class RefFrame:
    """One reference DXTImage (dxt_image.h:46) built from raw RGB."""

    def __init__(self, rgb: np.ndarray, is_intra: bool, search_area: int, err_threshold: int):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self.h, self.w = rgb.shape[:2]
        self._p = lib().mptc_ref_frame_new(self.w, self.h, rgb.ctypes.data, int(is_intra), search_area, err_threshold)
        self.nb = lib().mptc_ref_num_blocks(self._p)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().mptc_ref_frame_free(self._p)
            self._p = None

    def blocks(self) -> np.ndarray:
        out = np.empty(self.nb, dtype=np.uint64)
        lib().mptc_ref_get_blocks(self._p, out.ctypes.data)
        return out

    def reencode(self, prev: "RefFrame | None"):
        lib().mptc_ref_reencode(self._p, prev._p if prev is not None else None)

    def motion(self) -> np.ndarray:
        out = np.empty(2 * self.nb, dtype=np.uint8)
        lib().mptc_ref_get_motion(self._p, out.ctypes.data)
        return out

    def unique(self) -> np.ndarray:
        n = lib().mptc_ref_num_unique(self._p)
        out = np.empty(n, dtype=np.uint32)
        if n:
            lib().mptc_ref_get_unique(self._p, out.ctypes.data)
        return out

    def inter_pixel_search(self, prev: "RefFrame", search_area: int, defined: bool = True):
        """DXTImage::InterPixelSearch (dxt_image.cpp:776-832) for every block against `prev`.
        defined=True: the loop over the reference's own CompressedBlock methods with the candidate word
        built without the undefined behaviour of Get4X4InterpolationBlock (see oracle/ref_wrap.cpp);
        defined=False: the compiled function as it is (its results depend on stack garbage)."""
        nb = self.nb
        out = {"min_err": np.empty(nb, np.int32), "motion": np.empty(2 * nb, np.uint8), "index": np.empty(nb, np.uint32),
               "reassigned": np.empty(nb, np.uint8)}
        fn = lib().mptc_ref_inter_pixel_search_defined if defined else lib().mptc_ref_inter_pixel_search
        fn(self._p, prev._p, search_area, out["min_err"].ctypes.data, out["motion"].ctypes.data,
                                          out["index"].ctypes.data, out["reassigned"].ctypes.data)
        return out

    def psnr_logical(self) -> float:
        return lib().mptc_ref_psnr_logical(self._p)

    def psnr_physical(self) -> float:
        return lib().mptc_ref_psnr_physical(self._p)

    def entropy_payload(self) -> bytes:
        cap = 16 * self.nb + (1 << 20)
        buf = np.empty(cap, dtype=np.uint8)
        n = lib().mptc_ref_entropy_encode(self._p, buf.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("payload larger than buffer")
        return buf[:n].tobytes()

    def payload_planes(self, payload: bytes):
        """-> (n_unique, planes[6, nb] uint8, motion[2nb] uint8, sizes[5] uint32)."""
        src = np.frombuffer(payload, dtype=np.uint8)
        planes = np.empty(6 * self.nb, dtype=np.uint8)
        motion = np.empty(2 * self.nb, dtype=np.uint8)
        sizes = np.zeros(5, dtype=np.uint32)
        nu = lib().mptc_ref_payload_planes(src.ctypes.data, len(payload), self.nb, planes.ctypes.data,
                                           motion.ctypes.data, sizes.ctypes.data)
        if nu < 0:
            raise RuntimeError("payload did not parse")
        return nu, planes.reshape(6, self.nb), motion, sizes

    def times(self):
        t = np.zeros(3, dtype=np.float64)
        lib().mptc_ref_get_times(self._p, t.ctypes.data)
        return {"fit_s": t[0], "search_s": t[1], "entropy_s": t[2]}


def arith_encode(sym: np.ndarray) -> bytes:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    cap = 2 * sym.size + 1024
    out = np.empty(cap, dtype=np.uint8)
    n = lib().mptc_ref_arith_encode(sym.ctypes.data, sym.size, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].tobytes()


def selfcheck_png(rgb: np.ndarray) -> int:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    return lib().mptc_ref_selfcheck_png(rgb.shape[1], rgb.shape[0], rgb.ctypes.data)


def encode_sequence(frames: np.ndarray, search_area: int, err_threshold: int, gop: int):
    """Reference frame loop of CompressMultiUnique (codec.cpp:1383-1509) with
    intra_interval == unique_interval == gop.  Yields the RefFrame of each frame."""
    prev = None
    out = []
    for i, rgb in enumerate(frames):
        fr = RefFrame(rgb, i % gop == 0, search_area, err_threshold)
        fr.initial_blocks = fr.blocks()
        fr.reencode(prev)
        out.append(fr)
        prev = fr
    return out


def write_png(path: str, rgb: np.ndarray):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    if lib().mptc_ref_write_png(path.encode(), rgb.shape[1], rgb.shape[0], rgb.ctypes.data) != 0:
        raise RuntimeError("stbi_write_png failed")


def decode_stream(stream: bytes):
    """The reference's own decoder functions driven over a stream in memory (see
    mptc_ref_decode_stream): -> (blocks u64 [n, nb], rgb u8 [n, h, w, 3])."""
    import struct
    h, w, gop, _sa, n_groups = struct.unpack_from("<IIBBI", stream, 0)
    n = gop * n_groups
    nb = (w // 4) * (h // 4)
    buf = np.frombuffer(stream, dtype=np.uint8)
    blocks = np.empty((n, nb), dtype=np.uint64)
    rgb = np.empty((n, h, w, 3), dtype=np.uint8)
    lib().mptc_ref_decode_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    r = lib().mptc_ref_decode_stream(buf.ctypes.data, buf.size, blocks.ctypes.data, rgb.ctypes.data)
    if r != n:
        raise RuntimeError(f"reference decoder returned {r}")
    return blocks, rgb
