#!/usr/bin/env python
"""How much of the fp32 mode's drift is inherent?  CPU experiment with the fp64 oracle: all arithmetic in fp64, but the
STATE h, u, v is rounded to fp32 after every step (what any 24-bytes-per-update kernel must do).  Prints the normwise
distance from the pure fp64 run.  Result (profiles/r2_horizon_parity.md): ~1e-7 per step -- the same growth the fp32
kernels show, i.e. their fp32 ARITHMETIC adds little; the 1e-5 tolerance is a statement about <~ 100 steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sw_oracle as so  # noqa: E402


def run(M, N, nsteps, round_state, diffusion):
    og = so.OracleGrid(M, N)
    h, u, v, f = so.initial_conditions(0, og)
    h, u, v = (np.ascontiguousarray(x) for x in (h, u, v))
    if round_state:
        h, u, v = (np.ascontiguousarray(x.astype(np.float32).astype(np.float64)) for x in (h, u, v))
    stp = so.Stepper(og, f, None, diffusion)
    t, out = 0.0, {}
    for k in range(1, nsteps + 1):
        _, t = stp.step(h, u, v, 0.5, t, 1e12)
        if round_state:
            for q in (h, u, v):
                q[:] = q.astype(np.float32)
        if k in (1, 10, 25, 50, 100, 200):
            out[k] = (h.copy(), u.copy(), v.copy())
    return out


if __name__ == "__main__":
    for M, N, n, diff in ((360, 180, 200, False), (1440, 720, 100, True)):
        a, b = run(M, N, n, False, diff), run(M, N, n, True, diff)
        for k in a:
            e = [float(np.max(np.abs(x - y)) / np.max(np.abs(y))) for x, y in zip(b[k], a[k])]
            print(f"{M}x{N}{' + diffusion' if diff else ''}: step {k:3d}  h {e[0]:.1e}  u {e[1]:.1e}  v {e[2]:.1e}")
