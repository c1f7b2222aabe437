#!/usr/bin/env python
"""Host <-> device copy bandwidth of this box, as the e2e leg of bench.py sees it: pinned 1 GiB buffers,
current stream vs a side stream, and the process's CPU / NUMA placement (dev probe)."""
import os
import subprocess
import time

import torch

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
n = 1 << 27  # 1 GiB of float64
host = torch.empty(n, dtype=torch.float64).pin_memory()
host.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device=dev)
side = torch.cuda.Stream(dev)


def bw(fn, stream=None):
    best = 0.0
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if stream is None:
            fn()
        else:
            with torch.cuda.stream(stream):
                fn()
        torch.cuda.synchronize()
        best = max(best, n * 8 / (time.perf_counter() - t0) / 1e9)
    return best


print("affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)), "cpus")
print("H2D current stream  %.1f GB/s" % bw(lambda: d.copy_(host, non_blocking=True)))
print("H2D side stream     %.1f GB/s" % bw(lambda: d.copy_(host, non_blocking=True), side))
print("D2H current stream  %.1f GB/s" % bw(lambda: host.copy_(d, non_blocking=True)))
print("D2H side stream     %.1f GB/s" % bw(lambda: host.copy_(d, non_blocking=True), side))
print("H2D .to()           %.1f GB/s" % bw(lambda: host.to(dev)))
for cmd in (["nvidia-smi", "topo", "-m"], ["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name'"], ["bash", "-c", "numactl -s 2>/dev/null | head -5"]):
    try:
        print(subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout)
    except Exception as e:  # noqa: BLE001
        print(cmd, e)
