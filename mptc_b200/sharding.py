"""GOP sharding across GPUs/ranks: GOPs are independent, so each rank encodes a
contiguous range of GOPs with no data-path collective; the per-frame results are gathered on the
host of rank 0 and assembled in GOP order (what ThreadedCompressMultiUnique does with one thread
per group, /root/reference codec/codec.cpp:1781-1854).  torch.distributed is plumbing only."""
from __future__ import annotations

import numpy as np


def shard_gops(n_frames: int, gop: int, rank: int, world: int):
    """-> (first_frame, n_frames_of_this_rank): contiguous, GOP aligned, balanced to one GOP."""
    n_gops = (n_frames + gop - 1) // gop
    base, extra = divmod(n_gops, world)
    g0 = rank * base + min(rank, extra)
    g1 = g0 + base + (1 if rank < extra else 0)
    first = min(g0 * gop, n_frames)
    last = min(g1 * gop, n_frames)
    return first, last - first


def gather_results(local: dict, rank: int, world: int, group=None):
    """All ranks contribute {'motion','unique','n_unique','planes'} arrays of their own frames
    (first dimension = frames, possibly zero); rank 0 receives them concatenated in rank order,
    the other ranks get None."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    out = {}
    for key in ("motion", "unique", "n_unique", "planes"):
        t = torch.from_numpy(np.ascontiguousarray(local[key]))
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64), group=group)
        counts = [int(c.item()) for c in counts]
        if rank == 0:
            parts = [t]
            for r in range(1, world):
                buf = torch.empty((counts[r],) + tuple(t.shape[1:]), dtype=t.dtype)
                if counts[r]:
                    dist.recv(buf, src=r, group=group)
                parts.append(buf)
            out[key] = torch.cat(parts, dim=0).numpy()
        elif t.shape[0]:
            dist.send(t, dst=0, group=group)
    return out if rank == 0 else None
