import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sw_solver_b200.grid import LatLonGrid, CartesianGrid
from sw_solver_b200.ic import ICType, get_initial_conditions
from sw_solver_b200.solver import Stepper
M, N = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
ll = LatLonGrid(M, N); cg = CartesianGrid(ll)
h, u, v, f = get_initial_conditions(ICType.RossbyHaurwitzWave, ll)
st = Stepper(ll, cg, h, u, v, f, None, courant=0.5, final_time=1e9, use_diffusion=False, mode=mode)
st.enqueue(3)
s = st.status()
print("steps", s.steps, "time", s.time, "err", s.error)
os.environ["SWB_KERNEL"] = "ldg"
st2 = Stepper(ll, cg, h, u, v, f, None, courant=0.5, final_time=1e9, use_diffusion=False, mode=mode)
st2.enqueue(3)
a = st.to_numpy(); b = st2.to_numpy()
print("tma vs ldg max abs diff", [float(np.abs(x - y).max()) for x, y in zip(a, b)])
