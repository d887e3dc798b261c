"""Aggregate host<->device bandwidth of the box with all ranks copying at once (no kernels): the ceiling
for bench.py's end-to-end step, which moves 373 MB in and 137 MB out per rank and step.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/micro/pcie_aggregate.py"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
IN, OUT = 373_248_000, 136_765_988
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory()
h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step(both=True):
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    if both:
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)


for mode, both in (("H2D only", False), ("H2D + D2H concurrently", True)):
    for _ in range(3):
        step(both)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        step(both)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        bytes_ = IN + (OUT if both else 0)
        print(f"{world} ranks, {mode}: {t.item() * 1e3:.2f} ms per step (max over ranks) = {bytes_ / t.item() / 1e9:.1f} GB/s per rank, "
              f"{world * bytes_ / t.item() / 1e9:.1f} GB/s aggregate")
if world > 1:
    dist.destroy_process_group()
