"""CPU, world_size 2 over gloo: the GOP sharding + ordered gather + stream assembly path used for
N > 1 GPUs.  Per-frame results come from the oracle here (no GPU in this container); the stream
assembled from two shards must be byte-identical to the single-rank stream."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from mptc_b200.sharding import shard_gops


def test_shard_gops_covers_everything_once():
    for n_frames, gop in [(60, 15), (8, 4), (7, 3), (1, 5), (600, 15)]:
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, cnt = shard_gops(n_frames, gop, r, world)
                assert first % gop == 0 or cnt == 0
                seen += list(range(first, first + cnt))
            assert seen == list(range(n_frames))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mptc_b200 import capi
    from mptc_b200.sharding import gather_results, shard_gops
    from mptc_b200.synth import make_sequence
    from oracle import port as oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, n, sa, thr, gop = 256, 256, 4, 4, 50, 2
    frames = make_sequence(w, h, n, seed=77)
    nb = (w // 4) * (h // 4)
    first, cnt = shard_gops(n, gop, rank, world)
    local = {"motion": np.empty((cnt, 2 * nb), np.uint8), "unique": np.zeros((cnt, nb), np.uint32),
             "n_unique": np.zeros(cnt, np.uint32), "planes": np.empty((cnt, 6, h // 4, w // 4), np.uint8)}
    prev = None
    for k in range(cnt):
        f = first + k
        init = oracle.dxt1_fit(frames[f])
        blocks, mo, un = oracle.reencode(frames[f], f % gop == 0, sa, thr, init, prev)
        local["motion"][k] = mo
        local["unique"][k, : un.size] = un
        local["n_unique"][k] = un.size
        local["planes"][k] = oracle.endpoint_planes(blocks, w // 4, h // 4)
        prev = blocks
    res = gather_results(local, rank, world)
    if rank == 0:
        stream, st = capi.assemble_stream(w, h, sa, thr, gop, res["motion"], res["unique"], res["n_unique"], res["planes"], 2)
        q.put(stream)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_assembly_equals_reference_stream():
    import torch.multiprocessing as mp
    from golden_util import load
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stream = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stream == load("stream_256x256_sa4_gop2")["stream"].tobytes()
