import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sw_solver_b200.grid import CartesianGrid, LatLonGrid
from sw_solver_b200.ic import ICType, get_initial_conditions
from sw_solver_b200.solver import Stepper, band_layout, link_local

def run(M, N, world, nsteps, mode="fast", diff=False):
    ll = LatLonGrid(M, N); cg = CartesianGrid(ll)
    h, u, v, f = get_initial_conditions(ICType(0), ll)
    kw = dict(courant=0.5, final_time=1e9, use_diffusion=diff, mode=mode)
    one = Stepper(ll, cg, h, u, v, f, None, **kw); one.enqueue(nsteps); ref = one.to_numpy(); one.close()
    bands = band_layout(ll.N, world, 2 if diff else 1)
    sts = [Stepper(ll, cg, h, u, v, f, None, band=b, world=world, **kw) for b in bands]
    link_local(sts)
    for st in sts: st.cfl_init()
    for _ in range(nsteps):
        for st in sts: st.enqueue(1)
    for b, st in zip(bands, sts):
        loc = st.to_numpy(); s = st.status()
        for k in range(3):
            r = ref[k][:, b["j_begin"]:b["j_begin"] + b["n_local"]]
            bad = np.argwhere(loc[k] != r)
            if len(bad):
                print(f"world {world} steps {nsteps} rank {b['rank']} field {k}: {len(bad)} mismatches; i range {bad[:,0].min()}..{bad[:,0].max()}, "
                      f"jl range {bad[:,1].min()}..{bad[:,1].max()} (n_local {b['n_local']}, j_first-j_begin {b['j_first']-b['j_begin']}, j_last-j_begin {b['j_last']-b['j_begin']}) "
                      f"maxdiff {np.nanmax(np.abs(loc[k]-r)):.3e} steps {s.steps} err {s.error}; first {bad[:5].tolist()}")
        st.close()
    print("done", M, N, world, nsteps, os.environ.get("SWB_KERNEL"))

for world in (2, 8):
    for nsteps in (1, 2, 25):
        run(360, 180, world, nsteps)
