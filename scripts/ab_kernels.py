#!/usr/bin/env python
"""A/B of the step kernels on one GPU (dev tool): for each grid and each environment setting, the step time
over a burst (first `--steps` steps after warm-up) and, with --sustain S, over S seconds.

    python scripts/ab_kernels.py --grids 7200x3600,16384x8192 --envs "SWB_STREAM=0;SWB_STREAM=1" [--diffusion] [--dtype float32]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(a):
    sys.path.insert(0, ROOT)
    import torch

    from sw_solver_b200 import bands
    from sw_solver_b200.grid import CartesianGrid, LatLonGrid
    from sw_solver_b200.solver import Stepper

    M, N = (int(x) for x in a.grid.split("x"))
    dev = torch.device("cuda", 0)
    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    hd, ud, vd, f = (bands.zonal_flow_band if a.ic == "zgf" else bands.rossby_haurwitz_band)(ll, 0, ll.N, dev)
    if a.ic == "zgf" and os.environ.get("SWB_FSEP", "1") == "0":  # the 2-D array kernels: materialise f in the device layout
        f = torch.from_numpy(f.array().T.copy()).to(dev)
    st = Stepper(ll, cg, hd, ud, vd, f, None, courant=0.5, final_time=1e12, use_diffusion=a.diffusion, mode=a.mode, layout="device", dtype=a.dtype)
    del hd, ud, vd
    st.enqueue(a.warmup)
    torch.cuda.synchronize()

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.enqueue(n)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    l0 = st.launch_count
    burst = timed(a.steps)
    launches = st.launch_count - l0
    out = {"grid": a.grid, "env": a.tag, "kernel": "stream" if st.streaming else ("resident" if st.resident else "march"),
           "us_per_step_burst": 1e3 * burst, "launches": launches}
    if a.sustain > 0:
        n = max(a.steps, int(a.sustain * 1e3 / burst))
        out["us_per_step_sustained"] = 1e3 * timed(n)
        out["sustained_steps"] = n
    s = st.status()
    out["error"] = s.error
    pts = (M + 1) * (ll.N - 2)
    bpu = 48.0 if a.dtype == "float64" else 24.0
    out["frac_of_hbm_burst"] = bpu * pts / (burst * 1e-3) / 1e9 / 6553.9
    print(json.dumps(out), flush=True)
    st.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="7200x3600,16384x8192")
    ap.add_argument("--envs", default="SWB_STREAM=0;SWB_STREAM=1")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--sustain", type=float, default=0.0)
    ap.add_argument("--diffusion", action="store_true")
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--ic", default="rh", choices=["rh", "zgf"])
    ap.add_argument("--grid", default=None)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    if a.grid:
        return child(a)
    for g in a.grids.split(","):
        for env in a.envs.split(";"):
            e = dict(os.environ)
            for kv in env.split():
                if "=" in kv:
                    k, v = kv.split("=", 1)
                    e[k] = v
            cmd = [sys.executable, __file__, "--grid", g, "--tag", env, "--steps", str(a.steps), "--warmup", str(a.warmup), "--sustain", str(a.sustain),
                   "--mode", a.mode, "--dtype", a.dtype, "--ic", a.ic] + (["--diffusion"] if a.diffusion else [])
            r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=600)
            sys.stderr.write(r.stderr[-600:] if os.environ.get("SWB_VERBOSE") else ""); line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"grid": g, "env": env, "failed": r.stderr[-400:]})
            print(line, flush=True)


if __name__ == "__main__":
    main()
