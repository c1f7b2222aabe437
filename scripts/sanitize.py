"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sw_solver_b200.grid import CartesianGrid, LatLonGrid  # noqa: E402
from sw_solver_b200.ic import ICType, get_initial_conditions  # noqa: E402
from sw_solver_b200.solver import Stepper  # noqa: E402

CASES = [  # M, N, diffusion, mode, env
    (64, 48, True, "fast", {}),                                        # resident, direct loads, diffusion
    (64, 48, True, "strict", {}),
    (200, 60, False, "fast", {"SWB_KERNEL": "tma", "SWB_RES_CHUNK": "8"}),   # resident, TMA ring
    (200, 60, True, "fast", {"SWB_KERNEL": "tma", "SWB_RES_CHUNK": "8"}),    # resident, ring-fed diffusion star
    (200, 60, False, "fast", {"SWB_RESIDENT": "0", "SWB_KERNEL": "tma", "SWB_CHUNK": "8", "SWB_CFL_FILTER": "1"}),  # march, TMA, CFL filter
    (200, 60, True, "fast", {"SWB_RESIDENT": "0", "SWB_KERNEL": "tma", "SWB_CHUNK": "8"}),                          # march, ring-fed star
    (130, 40, True, "fast_guarded", {"SWB_RESIDENT": "0", "SWB_KERNEL": "ldg"}),
]
for M, N, diff, mode, env in CASES:
    for k in ("SWB_KERNEL", "SWB_RES_CHUNK", "SWB_RESIDENT", "SWB_CHUNK", "SWB_CFL_FILTER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    h, u, v, f = get_initial_conditions(ICType.RossbyHaurwitzWave, ll)
    st = Stepper(ll, cg, h, u, v, f, None, courant=0.5, final_time=1e9, use_diffusion=diff, mode=mode, device="cuda:0")
    st.enqueue(3)
    st.enqueue(4)
    s = st.status()
    hh, uu, vv = st.to_numpy()
    assert s.steps == 7 and s.error == 0 and np.isfinite(hh).all(), (M, N, mode, env)
    print(f"{M}x{N} diff={diff} {mode} {env}: resident={st.resident} ok", flush=True)
    st.close()
print("sanitize.py: all cases ran")
