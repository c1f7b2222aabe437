"""Step time of the L2-resident configurations (BASELINE.json configs 1 and 2 and friends):
launch-per-step `march_kernel` against the one-launch `resident_kernel`, with a bit-for-bit
comparison of the two results.  Prints one JSON line per case.

    python scripts/small_grids.py [--steps 2000] [--cases 360x180,1440x720d,...]

A case is MxN with an optional trailing "d" (diffusion on).  Environment knobs forwarded to
the library per run: SWB_RESIDENT, SWB_KERNEL, SWB_RES_CHUNK, SWB_CHUNK.
"""

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sw_solver_b200.grid import CartesianGrid, LatLonGrid  # noqa: E402
from sw_solver_b200.ic import ICType, get_initial_conditions  # noqa: E402
from sw_solver_b200.solver import Stepper  # noqa: E402


def run(M, N, diff, mode, steps, warmup, env):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        ll = LatLonGrid(M, N)
        cg = CartesianGrid(ll)
        h, u, v, f = get_initial_conditions(ICType.RossbyHaurwitzWave, ll)
        st = Stepper(ll, cg, h, u, v, f, None, courant=0.5, final_time=1e12, use_diffusion=diff, mode=mode, device="cuda:0")
    finally:
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val
    st.enqueue(warmup)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.enqueue(steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    s = st.status()
    state = [x.clone() for x in st.state()]
    res = dict(resident=st.resident, us_per_step=1e3 * best / steps, steps=int(s.steps), time=s.time, error=int(s.error))
    st.close()
    return res, state


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--cases", default="360x180,1440x720,1440x720d")
    ap.add_argument("--variants", default="march,resident", help="comma list of: march, resident, resident-ldg, resident-tma, resident-cN (chunk N)")
    args = ap.parse_args()
    for case in args.cases.split(","):
        diff = case.endswith("d")
        M, N = (int(x) for x in case.rstrip("d").split("x"))
        ref = None
        for var in args.variants.split(","):
            env = {"SWB_RESIDENT": 0 if var.split("-")[0] == "march" else 1}
            for tok in var.split("-")[1:]:
                if tok in ("ldg", "tma"):
                    env["SWB_KERNEL"] = tok
                elif tok.startswith("c"):
                    env["SWB_RES_CHUNK"] = int(tok[1:])
            try:
                res, state = run(M, N, diff, args.mode, args.steps, args.warmup, env)
            except Exception as e:  # a variant that does not apply to this case
                print(json.dumps({"case": case, "variant": var, "skipped": str(e)[:200]}), flush=True)
                continue
            if ref is None:
                ref = state
                same = None
            else:
                same = all(bool(torch.equal(a, b)) for a, b in zip(ref, state))
            pts = (M + 1) * (LatLonGrid(M, N).N - 2)
            res.update(case=case, variant=var, mode=args.mode, bit_identical_to_first=same,
                       gupdates_per_s=pts / res["us_per_step"] * 1e-3)
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
