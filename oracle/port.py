"""ctypes binding of oracle/libmptc_oracle.so (the plain-C restatement, mptc_oracle.c).
TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmptc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mptc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, ci = C.c_void_p, C.c_int
        L.mptc_oracle_dxt1_fit.argtypes = [vp, ci, ci, vp]
        L.mptc_oracle_reencode.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, vp, vp]
        L.mptc_oracle_eval_candidate.argtypes = [vp, C.c_uint64, C.c_uint32, vp, vp]
        L.mptc_oracle_endpoint_planes.argtypes = [vp, ci, ci, vp]
        L.mptc_oracle_arith_encode.argtypes = [vp, ci, vp, ci]
        L.mptc_oracle_arith_decode.argtypes = [vp, ci, vp, ci]
        L.mptc_oracle_inverse_planes.argtypes = [vp, ci, ci, vp, vp]
        L.mptc_oracle_decode_rgb.argtypes = [vp, ci, ci, vp]
        L.mptc_oracle_psnr.restype = C.c_double
        L.mptc_oracle_psnr.argtypes = [vp, ci, ci, vp]
        L.mptc_oracle_tables.argtypes = [vp, vp]
        L.mptc_oracle_reconstruct_words.argtypes = [vp, vp, ci, vp, ci, ci, ci, vp]
        L.mptc_oracle_check_blocks.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, ci]
        L.mptc_oracle_inter_pixel_search.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, vp]
        L.mptc_oracle_ips_pattern.argtypes = [ci, vp]
        L.mptc_oracle_encode_gops.restype = C.c_double
        L.mptc_oracle_encode_gops.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, vp, vp]
        _lib = L
    return _lib


def dxt1_fit(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    out = np.empty((h // 4) * (w // 4), dtype=np.uint64)
    lib().mptc_oracle_dxt1_fit(rgb.ctypes.data, w, h, out.ctypes.data)
    return out


def reencode(rgb, is_intra, search_area, err_threshold, init_blocks, prev_blocks=None):
    """-> (final_blocks u64[nb], motion u8[2nb], unique u32[n])."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    blocks = np.array(init_blocks, dtype=np.uint64, copy=True)
    nb = blocks.size
    motion = np.empty(2 * nb, dtype=np.uint8)
    unique = np.empty(nb, dtype=np.uint32)
    pp = None
    if prev_blocks is not None:
        prev_blocks = np.ascontiguousarray(prev_blocks, dtype=np.uint64)
        pp = prev_blocks.ctypes.data
    n = lib().mptc_oracle_reencode(rgb.ctypes.data, w, h, int(is_intra), search_area, err_threshold,
                                   blocks.ctypes.data, pp, motion.ctypes.data, unique.ctypes.data)
    return blocks, motion, unique[:n].copy()


def inter_pixel_search(rgb, search_area, cur_blocks, prev_blocks):
    """DXTImage::InterPixelSearch for every block -> dict(min_err i32[nb], motion u8[2nb], index u32[nb],
    reassigned u8[nb])."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    cur = np.ascontiguousarray(cur_blocks, dtype=np.uint64)
    prev = np.ascontiguousarray(prev_blocks, dtype=np.uint64)
    nb = cur.size
    out = {"min_err": np.empty(nb, np.int32), "motion": np.empty(2 * nb, np.uint8), "index": np.empty(nb, np.uint32),
           "reassigned": np.empty(nb, np.uint8)}
    lib().mptc_oracle_inter_pixel_search(rgb.ctypes.data, w, h, search_area, cur.ctypes.data, prev.ctypes.data,
                                         out["min_err"].ctypes.data, out["motion"].ctypes.data, out["index"].ctypes.data,
                                         out["reassigned"].ctypes.data)
    return out


def ips_pattern(search_area) -> np.ndarray:
    """DXTImage::SetPattern: (n, 2) int8 pixel offsets (i, j) in search order."""
    n = lib().mptc_oracle_ips_pattern(search_area, None)
    out = np.empty((n, 2), np.int8)
    lib().mptc_oracle_ips_pattern(search_area, out.ctypes.data)
    return out


def endpoint_planes(blocks, bw, bh) -> np.ndarray:
    """-> uint8 [6, pbh, pbw] with pbw/pbh = bw/bh rounded up to multiples of 64."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint64)
    pbw, pbh = (bw + 63) // 64 * 64, (bh + 63) // 64 * 64
    out = np.empty((6, pbh, pbw), dtype=np.uint8)
    lib().mptc_oracle_endpoint_planes(blocks.ctypes.data, bw, bh, out.ctypes.data)
    return out


def arith_encode(sym) -> bytes:
    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    cap = 2 * sym.size + 1024
    out = np.empty(cap, dtype=np.uint8)
    n = lib().mptc_oracle_arith_encode(sym.ctypes.data, sym.size, out.ctypes.data, cap)
    assert n >= 0
    return out[:n].tobytes()


def psnr(rgb, blocks) -> float:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    blocks = np.ascontiguousarray(blocks, dtype=np.uint64)
    return lib().mptc_oracle_psnr(rgb.ctypes.data, rgb.shape[1], rgb.shape[0], blocks.ctypes.data)


def tables():
    a = np.empty((256, 2), dtype=np.uint8)
    b = np.empty((256, 2), dtype=np.uint8)
    lib().mptc_oracle_tables(a.ctypes.data, b.ctypes.data)
    return a, b


def encode_gops(frames, gop, search_area, err_threshold, threads, want_outputs=True):
    """CPU baseline: -> (seconds, blocks[n, nb] or None, motion[n, 2nb] or None)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = frames.shape[:3]
    nb = (h // 4) * (w // 4)
    ob = np.empty((n, nb), dtype=np.uint64) if want_outputs else None
    om = np.empty((n, 2 * nb), dtype=np.uint8) if want_outputs else None
    t = lib().mptc_oracle_encode_gops(frames.ctypes.data, n, w, h, gop, search_area, err_threshold, threads,
                                      ob.ctypes.data if want_outputs else None,
                                      om.ctypes.data if want_outputs else None)
    return t, ob, om


def reconstruct_words(motion, unique, prev_words, bw, bh, search_area):
    """Decoder-side index reconstruction; -> (words u32[nb], n_unique_consumed)."""
    motion = np.ascontiguousarray(motion, dtype=np.uint8)
    unique = np.ascontiguousarray(unique, dtype=np.uint32)
    out = np.zeros(bw * bh, dtype=np.uint32)
    pw = None
    if prev_words is not None:
        prev_words = np.ascontiguousarray(prev_words, dtype=np.uint32)
        pw = prev_words.ctypes.data
    n = lib().mptc_oracle_reconstruct_words(motion.ctypes.data, unique.ctypes.data, unique.size, pw, bw, bh,
                                            search_area, out.ctypes.data)
    return out, n


def check_blocks(rgb, is_intra, search_area, err_threshold, init_blocks, cur_final, prev_final, motion, which):
    """Re-derives the reference's decision for the blocks in `which`; -> number of mismatches."""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    init_blocks = np.ascontiguousarray(init_blocks, dtype=np.uint64)
    cur_final = np.ascontiguousarray(cur_final, dtype=np.uint64)
    motion = np.ascontiguousarray(motion, dtype=np.uint8)
    which = np.ascontiguousarray(which, dtype=np.int32)
    pf = None
    if prev_final is not None:
        prev_final = np.ascontiguousarray(prev_final, dtype=np.uint64)
        pf = prev_final.ctypes.data
    return lib().mptc_oracle_check_blocks(rgb.ctypes.data, w, h, int(is_intra), search_area, err_threshold,
                                          init_blocks.ctypes.data, cur_final.ctypes.data, pf, motion.ctypes.data,
                                          which.ctypes.data, which.size)


def arith_decode(code: bytes, n: int) -> np.ndarray:
    """Arithmetic_Codec::decode x n with a fresh Adaptive_Data_Model(257) (codec.cpp:560-577)."""
    buf = np.zeros(len(code) + 8, dtype=np.uint8)   # the decoder reads ahead of the code
    buf[: len(code)] = np.frombuffer(code, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    r = lib().mptc_oracle_arith_decode(buf.ctypes.data, len(code), out.ctypes.data, n)
    if r != 0:
        raise ValueError("arithmetic code ran out")
    return out


def inverse_planes(planes: np.ndarray, bw: int, bh: int):
    """ReconstructEndPoints (codec.cpp:697-800): 6 symbol planes -> (ep1, ep2) RGB565 per block."""
    planes = np.ascontiguousarray(planes, dtype=np.uint8)
    ep1 = np.empty(bw * bh, dtype=np.uint16)
    ep2 = np.empty(bw * bh, dtype=np.uint16)
    lib().mptc_oracle_inverse_planes(planes.ctypes.data, bw, bh, ep1.ctypes.data, ep2.ctypes.data)
    return ep1, ep2


def decode_rgb(blocks, w: int, h: int) -> np.ndarray:
    """DXTImage::DecompressedImage of decoded physical blocks -> uint8 [h, w, 3]."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint64)
    out = np.empty((h, w, 3), dtype=np.uint8)
    lib().mptc_oracle_decode_rgb(blocks.ctypes.data, w, h, out.ctypes.data)
    return out


def parse_stream(stream: bytes):
    """Splits an MPTC stream (reader codec.cpp:1172-1184 + :1201-1290) into
    its header fields and per-frame compressed records.  Test helper, pure Python."""
    import struct
    h, w, gop, sa, n_groups = struct.unpack_from("<IIBBI", stream, 0)
    maxes = struct.unpack_from("<5I", stream, 14)
    off = 34
    groups = []
    for _ in range(n_groups):
        (cpal,) = struct.unpack_from("<I", stream, off); off += 4
        pal_code = stream[off:off + cpal]; off += cpal
        (unique_bytes,) = struct.unpack_from("<I", stream, off); off += 4
        frames = []
        for _ in range(gop):
            (n_unique,) = struct.unpack_from("<I", stream, off); off += 4
            recs = []
            for _ in range(5):
                (nb,) = struct.unpack_from("<I", stream, off); off += 4
                recs.append(stream[off:off + nb]); off += nb
            frames.append((n_unique, recs))
        groups.append((pal_code, unique_bytes, frames))
    assert off == len(stream)
    return {"h": h, "w": w, "gop": gop, "sa": sa, "n_groups": n_groups, "maxes": maxes, "groups": groups}


def decode_stream(stream: bytes) -> np.ndarray:
    """Whole-stream CPU decode built from the restated pieces (DecompressMultiUnique,
    codec.cpp:1161-1305): -> final 8-byte blocks [n_frames][nb]."""
    st = parse_stream(stream)
    w, h, gop, sa = st["w"], st["h"], st["gop"], st["sa"]
    bw, bh = w // 4, h // 4
    nb = bw * bh
    ps = ((bw + 63) // 64 * 64) * ((bh + 63) // 64 * 64)
    out = []
    for pal_code, unique_bytes, frames in st["groups"]:
        palette = arith_decode(pal_code, unique_bytes).view(np.uint32)
        at = 0
        prev_words = None
        for k, (n_unique, recs) in enumerate(frames):
            motion = arith_decode(recs[0], 2 * nb)
            planes = np.concatenate([arith_decode(recs[1], ps), arith_decode(recs[2], 2 * ps),
                                     arith_decode(recs[3], ps), arith_decode(recs[4], 2 * ps)])
            words, used = reconstruct_words(motion, palette[at:at + n_unique], prev_words if k else None, bw, bh, sa)
            assert used == n_unique
            at += n_unique
            ep1, ep2 = inverse_planes(planes, bw, bh)
            out.append(ep1.astype(np.uint64) | (ep2.astype(np.uint64) << np.uint64(16)) | (words.astype(np.uint64) << np.uint64(32)))
            prev_words = words
    return np.stack(out)
