#!/usr/bin/env python
"""A user script of ai2cm/sw_solver moved to the B200 path.

The reference's example (`examples/ex1-numpy.py` there) calls `sw_solver.numpy.solve(...)` and reads the snapshots out of
`save_data`.  Here the only change a user makes is the import: `sw_solver_b200` keeps the names (`ICType`, `LatLonGrid`,
`CartesianGrid`, `.numpy.solve`, the `save_data` keys and array shapes), the time loop runs on the GPU.

    python examples/rossby_wave_b200.py [--lon 30 --lat 30 --days 4 --mode fast|strict] [--check]

`--check` reruns the case in `mode="strict"` and reports the distance between the two (strict is bit-identical to the numpy
reference; fast must stay within 1e-12 normwise per field).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sw_solver_b200 as sw_solver  # noqa: E402  (the one line that differs from a reference script)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lon", type=int, default=30)
    ap.add_argument("--lat", type=int, default=30)
    ap.add_argument("--days", type=float, default=4.0)
    ap.add_argument("--courant", type=float, default=0.5)
    ap.add_argument("--diffusion", action="store_true")
    ap.add_argument("--interval", type=int, default=50, help="snapshot and print interval in steps")
    ap.add_argument("--mode", default="fast", choices=["fast", "fast_guarded", "strict"])
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()

    snapshots = {"interval": a.interval}
    h, u, v = sw_solver.numpy.solve(a.lon, a.lat, sw_solver.ICType.RossbyHaurwitzWave, a.days, a.courant, a.diffusion,
                                    save_data=snapshots, print_interval=a.interval, mode=a.mode)
    t = snapshots["t"]
    print(f"final state {h.shape}, {len(t)} snapshots up to t = {t[-1] / 3600.0:.2f} h; "
          f"h in [{h.min():.1f}, {h.max():.1f}] m, max |u| {np.abs(u).max():.2f} m/s, max |v| {np.abs(v).max():.2f} m/s")
    for key in ("h", "u", "v", "phi", "theta"):
        print(f"  save_data[{key!r}]: {snapshots[key].shape} {snapshots[key].dtype}")
    if a.check:
        ref = {"interval": a.interval}
        hs, us, vs = sw_solver.numpy.solve(a.lon, a.lat, sw_solver.ICType.RossbyHaurwitzWave, a.days, a.courant, a.diffusion,
                                           save_data=ref, mode="strict")
        worst = max(np.abs(x - y).max() / np.abs(y).max() for x, y in ((h, hs), (u, us), (v, vs)))
        print(f"  {a.mode} against strict: {worst:.2e} normwise (tolerance 1e-12), same snapshot times: {np.array_equal(t, ref['t']) or np.allclose(t, ref['t'], rtol=1e-13)}")
        return 0 if worst < 1e-12 else 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
