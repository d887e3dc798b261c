"""Structure of the blocks the inter search leaves to the intra wavefront (bench workload)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mptc_b200 import capi
from mptc_b200.synth import make_frame
W, H, SA, THR, GOP = 1920, 1080, 16, 50, 15
bw, bh = W // 4, H // 4
frames = np.stack([make_frame(W, H, f) for f in range(GOP)])
ctx = capi.Context(0)
out = ctx.encode_sequence(frames, SA, THR, GOP)
for f in range(GOP):
    m = out["motion"][f].reshape(bh, bw, 2)
    uniq = (m[..., 0] == 255) & (m[..., 1] == 255)
    inter = (m[..., 0] >= 128) & (m[..., 1] >= 128) & ~uniq
    left = ~inter
    intra = left & ~uniq
    rows = left.any(axis=1)
    g = left.reshape(bh, bw // 32, 32).sum(axis=2)
    first = np.where(rows, left.argmax(axis=1), bw)
    print(f"frame {f:2d}: leftover {left.sum():6d} ({100*left.mean():5.1f}%) intra-found {intra.sum():6d} unique {uniq.sum():5d} "
          f"rows {rows.sum():3d} groups {int((g>0).sum()):4d} sparse(<=2) {int(((g>0)&(g<=2)).sum()):4d} dense {int((g>2).sum()):4d} "
          f"max/row {left.sum(axis=1).max():3d} words distinct {len(np.unique(out['blocks'][f] >> 32)):6d}")
