import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sw_solver_b200.ic import ICType
from sw_solver_b200.solver import solve
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_data_10.npz"))
def nw(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for mode, dtype in (("fast", "float64"), ("fast", "float32"), ("fast_guarded", "float32")):
    sd = {"interval": 392}
    solve(10, 10, ICType.RossbyHaurwitzWave, 2, 0.5, True, save_data=sd, mode=mode, dtype=dtype)
    print(mode, dtype, {k: f"{nw(sd[k], g[k]):.2e}" for k in ("h", "u", "v")}, "t", sd["t"], "allclose", {k: bool(np.allclose(sd[k], g[k])) for k in ("h", "u", "v")})
