#!/usr/bin/env python
"""Error growth of the benched arithmetic (FAST fp64, fp32) against STRICT (bit-exact to numpy on
every grid the oracle reaches) over the bench horizon of BASELINE configs 3-5, on one GPU.

    python scripts/horizon_parity.py [--grids 7200x3600,16384x8192] [--steps 200,110] [--out profiles/x.json]

Prints one line per checkpoint: normwise error (max|diff| / max|ref|) of h, u, v and of the dt
history.  Everything stays on the device (two Steppers side by side, torch reductions).
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sw_solver_b200.bands import rossby_haurwitz_band  # noqa: E402
from sw_solver_b200.grid import CartesianGrid, LatLonGrid  # noqa: E402
from sw_solver_b200.solver import Stepper  # noqa: E402


def make(ll, cg, mode, dtype, diff, cap):
    hd, ud, vd, f_lat = rossby_haurwitz_band(ll, 0, ll.N, "cuda")
    st = Stepper(ll, cg, hd, ud, vd, f_lat, None, courant=0.5, final_time=1e12, use_diffusion=diff, mode=mode,
                 layout="device", log_capacity=cap, dtype=dtype)
    return st


def normwise_dev(a, b):
    den = float(b.abs().max())
    return float((a.double() - b).abs().max()) / (den if den > 0 else 1.0)


def run(M, N, nsteps, diff, checkpoints):
    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    ref = make(ll, cg, "strict", "float64", diff, nsteps)
    arms = {"fast": make(ll, cg, "fast", "float64", diff, nsteps), "fp32": make(ll, cg, "fast", "float32", diff, nsteps)}
    rows = []
    done = 0
    for cp in checkpoints:
        for st in (ref, *arms.values()):
            st.enqueue(cp - done)
        done = cp
        s = ref.status()
        assert s.steps == cp and s.error == 0, (s.steps, s.error)
        rq = ref._raw[s.current, :, :, : M + 3]
        rdt = ref.read_log(cp)[0]
        for name, st in arms.items():
            sa = st.status()
            assert sa.steps == cp and sa.error == 0, (name, sa.steps, sa.error)
            q = st._raw[sa.current, :, :, : M + 3]
            errs = [normwise_dev(q[k], rq[k]) for k in range(3)]
            dts = st.read_log(cp)[0]
            edt = float(np.max(np.abs(dts - rdt)) / np.max(np.abs(rdt)))
            row = {"grid": f"{M}x{ll.N}", "diffusion": diff, "arm": name, "steps": cp, "h": errs[0], "u": errs[1], "v": errs[2], "dt": edt,
                   "time_rel": abs(sa.time - s.time) / s.time}
            rows.append(row)
            print(json.dumps(row), flush=True)
    for st in (ref, *arms.values()):
        st.close()
    del ref, arms
    torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="1440x720d,7200x3600,16384x8192")
    ap.add_argument("--steps", default="100,200,110")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = []
    for g, n in zip(a.grids.split(","), a.steps.split(",")):
        diff = g.endswith("d")
        M, N = (int(x) for x in g.rstrip("d").split("x"))
        n = int(n)
        cps = sorted({c for c in (1, 2, 10, 25, 50, 100, 110, 200, 400, n) if c <= n})
        rows += run(M, N, n, diff, cps)
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
