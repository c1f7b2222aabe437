#!/usr/bin/env python
"""Where `solve(..., save_data={"interval": 10})` spends its wall time (dev probe): cProfile of the reference's call on the 1-degree grid."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sw_solver_b200.ic import ICType  # noqa: E402
from sw_solver_b200.solver import solve  # noqa: E402

solve(360, 180, ICType.RossbyHaurwitzWave, 0.01, 0.5, False)  # warm-up (library, context)
for sd in (None, {"interval": 10}):
    t0 = time.perf_counter()
    solve(360, 180, ICType.RossbyHaurwitzWave, 1.0, 0.5, False, save_data=sd)
    print("save_data", sd if sd is None else "interval 10", "wall %.3f s" % (time.perf_counter() - t0), flush=True)
sd = {"interval": 10}
pr = cProfile.Profile()
pr.enable()
solve(360, 180, ICType.RossbyHaurwitzWave, 1.0, 0.5, False, save_data=sd)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
