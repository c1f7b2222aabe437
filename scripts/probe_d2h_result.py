#!/usr/bin/env python
"""How to bring a 388 MB result array (742 stacked snapshots of the 1-degree grid) to the host fastest (dev probe)."""
import time

import torch

dev = torch.device("cuda", 0)
x = torch.randn(361, 180, 743, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
nb = x.numel() * 8


def t(name, fn, reps=3):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); y = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:34s} rep {r}: {1e3 * dt:7.1f} ms  {nb / dt / 1e9:6.1f} GB/s", flush=True)
        del y


t(".cpu()", lambda: x.cpu())
t("torch.zeros + copy_", lambda: torch.zeros(x.shape, dtype=x.dtype).copy_(x))
t("torch.empty + copy_", lambda: torch.empty(x.shape, dtype=x.dtype).copy_(x))
t("empty(pin_memory=True) + copy_", lambda: torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x, non_blocking=True))
print("threads", torch.get_num_threads())
