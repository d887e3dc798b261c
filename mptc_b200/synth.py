"""Deterministic, integer-only synthetic RGB frame generator.

The reference ships no test content and reads PNG directories
(/root/reference/codec/codec.cpp:1323-1328); benchmarks and parity tests here feed raw
RGB8 frames (row-major, stride 3*W) produced by this generator to BOTH the oracle and
the CUDA path.  Everything is uint32/int32 arithmetic so any language reproduces it.

Content per frame f:
  * smooth moving gradients + +-2 hash noise (most blocks are "found" by the search,
    which keeps the reference's 1 MiB entropy buffers from overflowing, codec.cpp:82);
  * three 64x64 textured sprites moving 8 px/frame (exercise the inter search);
  * one static flat rectangle (constant colour -> all-equal index words, the den==0
    path of RecalculateEndpoints, dxt_image.cpp:320);
  * a black/white checker pair (high contrast -> unique blocks).
"""
from __future__ import annotations

import numpy as np

__all__ = ["lowbias32", "make_frame", "make_sequence"]


def lowbias32(x: np.ndarray) -> np.ndarray:
    """32-bit integer hash (public-domain 'lowbias32' constants)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def make_frame(w: int, h: int, f: int, seed: int = 1234) -> np.ndarray:
    """Returns frame ``f`` as a C-contiguous uint8 array of shape (h, w, 3)."""
    if w % 4 or h % 4:
        raise ValueError("width/height must be multiples of 4 (dxt_image.cpp:428 reads OOB otherwise)")
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    with np.errstate(over="ignore"):
        key = (x + 65537 * y + 2654435761 * f + seed) & 0xFFFFFFFF
        hs = lowbias32(key.astype(np.uint32)).astype(np.int64)
    n1 = (hs % 5) - 2
    n2 = ((hs >> 8) % 5) - 2
    n3 = ((hs >> 16) % 5) - 2
    r = (x // 4 + 4 * f + 0 * y) % 256 + n1  # '0 * y' only broadcasts to (h, w)
    g = (y // 2 + x // 8 + 4 * f) % 256 + n2
    b = ((x * y) >> 12) % 256 + 8 * f + n3
    img = np.stack([r, g, b], axis=-1)
    img = np.clip(img, 0, 255)

    def put(x0, y0, sw, sh, patch):
        x0 = int(x0) % max(1, (w - sw + 1))
        y0 = int(y0) % max(1, (h - sh + 1))
        img[y0:y0 + sh, x0:x0 + sw, :] = patch[: min(sh, h - y0), : min(sw, w - x0), :]

    s = min(64, w // 4, h // 4)
    s -= s % 4
    if s >= 8:
        sx = np.arange(s, dtype=np.int64)[None, :]
        sy = np.arange(s, dtype=np.int64)[:, None]
        # three textured sprites, moving +8 px/frame in different directions
        t0 = np.stack([(sx * 4) % 256 + 0 * sy, (sy * 4) % 256 + 0 * sx, ((sx ^ sy) * 8) % 256], -1)
        t1 = np.stack([((sx + sy) * 3) % 256, ((sx * sy) >> 2) % 256, (255 - sx * 3) % 256 + 0 * sy], -1)
        t2 = np.stack([((sx // 8 + sy // 8) % 2) * 200 + 20, ((sx // 4) % 2) * 120 + 60 + 0 * sy,
                       (sy * 2) % 256 + 0 * sx], -1)
        put(w // 8 + 8 * f, h // 8, s, s, t0)
        put(w // 2, h // 8 + 8 * f, s, s, t1)
        put(w // 3 + 8 * f, h // 2 + 8 * f, s, s, t2)
        # static flat rectangle (block aligned so whole blocks are constant)
        fx = (3 * w // 4) // 4 * 4
        fy = (3 * h // 4) // 4 * 4
        fw = min(2 * s, w - fx)
        fh = min(s, h - fy)
        img[fy:fy + fh, fx:fx + fw, :] = np.array([90, 140, 200], dtype=np.int64)
        # black/white pair
        cx = (w // 16) // 4 * 4
        cy = (5 * h // 8) // 4 * 4
        cw = min(s, w - cx)
        ch = min(s, h - cy)
        chk = (((sx // 2) + (sy // 2)) % 2 * 255)[:ch, :cw]
        img[cy:cy + ch, cx:cx + cw, :] = chk[..., None]
    return np.ascontiguousarray(img.astype(np.uint8))


def make_sequence(w: int, h: int, n_frames: int, seed: int = 1234, start: int = 0) -> np.ndarray:
    """(n_frames, h, w, 3) uint8."""
    return np.stack([make_frame(w, h, start + f, seed) for f in range(n_frames)], axis=0)
