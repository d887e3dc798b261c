"""Write-only and copy bandwidth of the box (torch), as a yardstick for k_dxt1_to_rgb whose traffic
is 86 % writes: fill_ of 1 GiB (write only) vs copy_ of 1 GiB (read + write), best of 10."""
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn, bytes_ in (("fill (write only)", lambda: a.fill_(7), n), ("copy (read + write)", lambda: b.copy_(a), 2 * n)):
    best = 1e9
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:22s} {bytes_ / best / 1e6:8.1f} GB/s")
