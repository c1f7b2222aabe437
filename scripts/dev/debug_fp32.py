import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sw_solver_b200.grid import CartesianGrid, LatLonGrid
from sw_solver_b200.solver import Stepper
from oracle import sw_oracle as so

def nw(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for (M, N, diff) in ((360, 180, False), (1440, 720, False)):
    og = so.OracleGrid(M, N)
    h, u, v, f = so.initial_conditions(0, og)
    h, u, v = (np.ascontiguousarray(x) for x in (h, u, v))
    ll = LatLonGrid(M, N); cg = CartesianGrid(ll)
    st = Stepper(ll, cg, h, u, v, f, None, courant=0.5, final_time=1e9, use_diffusion=diff, mode="fast", dtype="float32")
    stp = so.Stepper(og, f, None, diff)
    t = 0.0; done = 0
    for target in (1, 10, 30, 100, 300):
        for _ in range(target - done):
            _, t = stp.step(h, u, v, 0.5, t, 1e9)
        st.enqueue(target - done); done = target
        hg, ug, vg = st.to_numpy()
        k = np.unravel_index(np.argmax(np.abs(vg - v)), v.shape)
        ku = np.unravel_index(np.argmax(np.abs(ug - u)), u.shape)
        print(f"{M}x{N} steps {target}: h {nw(hg,h):.2e} u {nw(ug,u):.2e} v {nw(vg,v):.2e}; worst v at {k} (dv {vg[k]-v[k]:.2e}, v {v[k]:.3f}), worst u at {ku} (du {ug[ku]-u[ku]:.2e}); max|u| {np.abs(u).max():.1f} max|v| {np.abs(v).max():.1f}")
    st.close()
