#!/usr/bin/env python
"""Where the end-to-end leg of bench.py spends its upload phase on a one-GPU box (dev probe): plain pinned copies of one
field, then the pieces of `Stepper(host h, u, v)` one by one -- allocation, context, upload -- cold and warm."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sw_solver_b200.grid import CartesianGrid, LatLonGrid  # noqa: E402
from sw_solver_b200.ic import coriolis_latitude, rossby_haurwitz_window  # noqa: E402
from sw_solver_b200.solver import Stepper  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
M, N = 16384, 8192
ll = LatLonGrid(M, N)
cg = CartesianGrid(ll)
f_lat = coriolis_latitude(ll)
print("cpus", len(os.sched_getaffinity(0)), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads(), flush=True)
t0 = time.perf_counter()
host = [torch.from_numpy(q).pin_memory() for q in rossby_haurwitz_window(ll, 0, N)]
out = [torch.empty_like(q).pin_memory() for q in host]
print("pinned host buffers: %.2f s" % (time.perf_counter() - t0), flush=True)
nb = host[0].numel() * 8


def sync():
    torch.cuda.synchronize(dev)


d = torch.empty_like(host[0], device=dev)
for rep in range(3):
    sync(); t0 = time.perf_counter(); d.copy_(host[0], non_blocking=True); sync(); t1 = time.perf_counter()
    out[0].copy_(d, non_blocking=True); sync(); t2 = time.perf_counter()
    print("plain copy of one field: H2D %.1f GB/s, D2H %.1f GB/s" % (nb / (t1 - t0) / 1e9, nb / (t2 - t1) / 1e9), flush=True)
# both directions at once (two streams)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
d2 = torch.empty_like(d)
sync(); t0 = time.perf_counter()
with torch.cuda.stream(s1):
    d.copy_(host[1], non_blocking=True)
with torch.cuda.stream(s2):
    out[2].copy_(d2, non_blocking=True)
sync(); t1 = time.perf_counter()
print("both directions at once: %.1f GB/s each" % (nb / (t1 - t0) / 1e9), flush=True)
del d, d2
torch.cuda.empty_cache()

kw = dict(courant=0.5, final_time=1e12, use_diffusion=False, mode="fast", device=dev)
for rep in range(3):
    if rep == 1:
        torch.cuda.empty_cache()
    sync(); t0 = time.perf_counter()
    z = torch.zeros((2, 3, N, 16448), dtype=torch.float64, device=dev); sync(); t1 = time.perf_counter()
    del z
    if rep == 1:
        torch.cuda.empty_cache()
    sync(); t2 = time.perf_counter()
    st = Stepper(ll, cg, host[0], host[1], host[2], f_lat, None, **kw); sync(); t3 = time.perf_counter()
    st.enqueue(20); cur = st.status().current; t4 = time.perf_counter()
    st.download(cur, out); t5 = time.perf_counter()
    st.reload(host[0], host[1], host[2]); sync(); t6 = time.perf_counter()
    st.close(); del st; t7 = time.perf_counter()
    print("rep %d: zeros(6.4 GB) %.1f ms | Stepper(host) %.1f ms (%.1f GB/s) | 20 steps %.1f ms | download %.1f ms (%.1f GB/s) | reload %.1f ms (%.1f GB/s) | close %.1f ms"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t3 - t2), 3 * nb / (t3 - t2) / 1e9, 1e3 * (t4 - t3), 1e3 * (t5 - t4), 3 * nb / (t5 - t4) / 1e9,
             1e3 * (t6 - t5), 3 * nb / (t6 - t5) / 1e9, 1e3 * (t7 - t6)), flush=True)
