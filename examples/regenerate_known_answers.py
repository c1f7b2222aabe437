#!/usr/bin/env python
"""Regenerates the reference's known-answer case on the B200 path and compares it with the reference's own files.

The reference ships `data/swes-numpy-ref-10-{h,u,v,t,phi,theta}.npy` (10 x 10 grid, Rossby-Haurwitz wave, 2 days, CFL 0.5, diffusion on,
one snapshot after 392 steps) and checks its solver against them with `np.allclose` (`tests/numpy_test.py:9-30`).  This script runs
the same case through `sw_solver_b200` (`mode="strict"`: numpy's arithmetic operation for operation), writes `swes-b200-ref-10-*.npy`
into `--out`, and compares with the copies of the reference's arrays kept in `tests/golden/ref_data_10.npz`.

    python examples/regenerate_known_answers.py [--out /tmp/known] [--mode strict|fast]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import sw_solver_b200 as swb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--mode", default="strict", choices=["strict", "fast", "fast_guarded"])
    a = ap.parse_args()
    case = dict(lon=10, lat=10, days=2, courant=0.5, diffusion=True, interval=392)
    save = {"interval": case["interval"]}
    swb.numpy.solve(case["lon"], case["lat"], swb.ICType.RossbyHaurwitzWave, case["days"], case["courant"], case["diffusion"],
                    save_data=save, mode=a.mode)
    if a.out:
        os.makedirs(a.out, exist_ok=True)
        for key in ("h", "u", "v", "t", "phi", "theta"):
            np.save(os.path.join(a.out, f"swes-b200-ref-10-{key}.npy"), save[key])
    golden = np.load(os.path.join(ROOT, "tests", "golden", "ref_data_10.npz"))
    ok = True
    for key in ("h", "u", "v", "t", "phi", "theta"):
        want = golden[key]
        got = np.asarray(save[key])
        close = got.shape == want.shape and np.allclose(got, want, rtol=1e-5, atol=1e-8)  # the reference's own criterion
        err = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)) if got.shape == want.shape else float("nan")
        ok &= bool(close)
        print(f"{key:6s} {str(got.shape):16s} allclose {close}  normwise {err:.2e}")
    print("matches the reference's known answers" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
