"""The UNMODIFIED reference numpy solver, staged next to the oracle -- TEST / BENCH INFRASTRUCTURE.

The reference's numpy path is four pure-Python modules (``src/sw_solver/{utils,grid,ic,numpy}.py``;
its ``__init__`` imports gt4py, which is not installed, so the package cannot be imported as
shipped).  ``stage()`` -- called by ``__graft_entry__.build()`` in the build container, where
``/root/reference`` exists -- copies those four files verbatim into the git-ignored directory
``oracle/_ref/sw_solver/``; like the built shared objects they then travel to the GPU box with the
snapshot (``oracle/_ref/`` is not gpurun-ignored).  Nothing of the reference enters the repository's
history, and nothing under ``sw_solver_b200/`` ever imports this module.

Uses: ``bench.py`` times the reference's own ``solve`` on the GPU box's host cores (SURVEY 8d "CPU
reference beside it", numpy.py:409-575); ``tests/test_oracle.py`` checks the C oracle against it live.
"""

import importlib.util
import os
import shutil
import sys
import time
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
PKG_DIR = os.path.join(REF_DIR, "sw_solver")
MODULES = ("utils", "grid", "ic", "numpy")
SOURCE = os.environ.get("SW_REFERENCE", "/root/reference")


def stage(force: bool = False) -> bool:
    """Copy the reference's four numpy-path modules into oracle/_ref/ (build container only).
    Returns True when the staged copy exists afterwards."""
    src_dir = os.path.join(SOURCE, "src", "sw_solver")
    if os.path.isdir(src_dir):
        os.makedirs(PKG_DIR, exist_ok=True)
        for name in MODULES:
            src, dst = os.path.join(src_dir, f"{name}.py"), os.path.join(PKG_DIR, f"{name}.py")
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                shutil.copyfile(src, dst)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(PKG_DIR, f"{n}.py")) for n in MODULES)


_mods = None


def load():
    """Import the staged modules under a private stub package (the reference's own relative imports
    resolve inside it; its gt4py-importing ``__init__`` is never executed)."""
    global _mods
    if _mods is None:
        if not available():
            raise ImportError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
        pkg_name = "_sw_solver_ref"
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [PKG_DIR]
        sys.modules[pkg_name] = pkg
        mods = {}
        for name in MODULES:
            spec = importlib.util.spec_from_file_location(f"{pkg_name}.{name}", os.path.join(PKG_DIR, f"{name}.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"{pkg_name}.{name}"] = mod
            spec.loader.exec_module(mod)
            setattr(pkg, name, mod)
            mods[name] = mod
        _mods = mods
    return _mods


def solve_steps(M, N, ic_type, nsteps, courant=0.5, use_diffusion=True, dts=None):
    """Run the reference's ``solve`` for exactly ``nsteps`` time steps: the final time is put half a
    step before the end of step ``nsteps`` of the dt history ``dts`` (from the oracle, which is pinned
    bit-exact to the reference, dt history included), so the loop's own clamp ends the run there.
    Returns (h, u, v, final_time, wall_seconds)."""
    ref = load()
    if dts is None or len(dts) < nsteps:
        from . import sw_oracle as so

        sd = {}
        so.solve(M, N, int(ic_type), 1e9, courant, use_diffusion, save_data=sd, max_steps=nsteps)
        dts = sd["_dts"]
    final_time = float(sum(dts[: nsteps - 1]) + 0.5 * dts[nsteps - 1]) if nsteps > 0 else 0.0
    sim_days = final_time / 86400.0
    t0 = time.perf_counter()
    h, u, v = ref["numpy"].solve(M, N, ref["ic"].ICType(int(ic_type)), sim_days, courant, use_diffusion)
    wall = time.perf_counter() - t0
    return h, u, v, 86400.0 * sim_days, wall


def time_loop_body(M, N, k_small, k_large, use_diffusion, courant=0.5):
    """Seconds per iteration of numpy.py:496-549 on an M x N grid: two unmodified ``solve`` runs of
    k_small and k_large steps, differenced (grid set-up, initial condition and diffusion coefficients
    cancel).  Returns a dict with the per-step time and updates/s."""
    from . import sw_oracle as so

    sd = {}
    so.solve(M, N, 0, 1e9, courant, use_diffusion, save_data=sd, max_steps=k_large)
    dts = sd["_dts"]
    *_, wa = solve_steps(M, N, 0, k_small, courant, use_diffusion, dts)
    *_, wb = solve_steps(M, N, 0, k_large, courant, use_diffusion, dts)
    per_step = (wb - wa) / (k_large - k_small)
    n_even = N + N % 2
    pts = (M + 1) * (n_even - 2)
    return {"grid": f"{M}x{n_even}", "diffusion": bool(use_diffusion), "steps": k_large - k_small, "seconds_per_step": per_step,
            "value": pts / per_step, "unit": "updates/s", "cores": 1, "wall_small": wa, "wall_large": wb, "pts_per_step": pts}
