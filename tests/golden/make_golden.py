"""Generate the golden fixtures in this directory from the *reference* numpy solver.

Run in the build container only (``/root/reference`` is absent on the GPU box)::

    python tests/golden/make_golden.py

The reference package cannot be imported as shipped (``src/sw_solver/__init__.py:3``
imports gt4py, which is not installed), so its four numpy-only modules
(``utils``, ``grid``, ``ic``, ``numpy``) are loaded under a stub ``sw_solver`` package.
Nothing from the reference is copied: only arrays it *computes* are stored.

Fixtures written
----------------
``ref_data_10.npz``     the six ``data/swes-numpy-ref-10-*.npy`` known-answer arrays
                        (``tests/numpy_test.py:9-30`` of the reference).
``grid_ic.npz``         LatLonGrid / CartesianGrid attributes and both ICs on a few grids.
``single_step.npz``     one ``lax_wendroff_update`` (+ diffusion) at dt=1.25 on the grids the
                        reference tests use ((4,4), (5,6)) plus a few more, both ICs,
                        and seeded random states (including negative-depth cells and
                        non-flat topography).
``multi_step.npz``      short full ``solve`` runs (both ICs, with/without diffusion) with
                        every step saved, plus the dt/time sequence.
"""

import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("SW_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """Import the reference's numpy path under a stub package (skipping __init__)."""
    pkg = types.ModuleType("sw_solver")
    pkg.__path__ = [os.path.join(REF, "src", "sw_solver")]
    sys.modules["sw_solver"] = pkg
    mods = {}
    for name in ("utils", "grid", "ic", "numpy"):
        spec = importlib.util.spec_from_file_location(
            f"sw_solver.{name}", os.path.join(REF, "src", "sw_solver", f"{name}.py")
        )
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"sw_solver.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods


def grid_ic_fixture(ref):
    out = {}
    sizes = [(4, 4), (5, 6), (10, 10), (7, 9), (37, 24), (64, 32)]
    out["sizes"] = np.asarray(sizes)
    for (M, N) in sizes:
        ll = ref["grid"].LatLonGrid(M, N)
        cg = ref["grid"].CartesianGrid(ll)
        tag = f"{M}x{N}"
        for name in ("phi", "theta", "c", "cMidy", "tg", "tgMidx", "tgMidy"):
            out[f"{tag}/ll/{name}"] = getattr(ll, name)
        for name in ("x", "y", "dx", "dy", "dy1", "dxc", "dyc", "dy1c"):
            out[f"{tag}/cg/{name}"] = getattr(cg, name)
        out[f"{tag}/cg/dxmin"] = np.float64(cg.dxmin)
        out[f"{tag}/cg/dymin"] = np.float64(cg.dymin)
        dc = ref["numpy"].DiffusionCoefficients(cg)
        for name in ("Ax", "Bx", "Cx", "Ay", "By", "Cy"):
            out[f"{tag}/dc/{name}"] = getattr(dc, name)
        for ic in ref["ic"].ICType:
            h, u, v, f = ref["ic"].get_initial_conditions(ic, ll)
            for name, arr in zip("huvf", (h, u, v, f)):
                out[f"{tag}/ic{int(ic)}/{name}"] = arr
    return out


def one_step(ref, ll, cg, dt, f, hs, h, u, v, diffusion):
    """LW update (+ diffusion) exactly as the reference loop body applies it
    (numpy.py:536-544), returning the (M+1, N-2) interior arrays."""
    nu = ref["utils"].EARTH_CONSTANTS.nu
    hn, un, vn = ref["numpy"].lax_wendroff_update(ll, cg, dt, f, hs, h, u, v)
    if diffusion:
        dc = ref["numpy"].DiffusionCoefficients(cg)
        hn += dt * nu * ref["numpy"].compute_diffusion(h, dc)
        un += dt * nu * ref["numpy"].compute_diffusion(u, dc)
        vn += dt * nu * ref["numpy"].compute_diffusion(v, dc)
    return hn, un, vn


def single_step_fixture(ref):
    out = {}
    cases = []
    dt = 1.25
    for (M, N) in [(4, 4), (5, 6), (10, 10), (7, 9), (12, 8), (37, 24)]:
        ll = ref["grid"].LatLonGrid(M, N)
        cg = ref["grid"].CartesianGrid(ll)
        shape = ll.shape
        for ic in ref["ic"].ICType:
            h, u, v, f = ref["ic"].get_initial_conditions(ic, ll)
            hs = np.zeros(shape)
            for diff in (False, True):
                tag = f"{M}x{N}/ic{int(ic)}/d{int(diff)}"
                hn, un, vn = one_step(ref, ll, cg, dt, f, hs, h, u, v, diff)
                out[f"{tag}/hnew"], out[f"{tag}/unew"], out[f"{tag}/vnew"] = hn, un, vn
                cases.append(tag)
        # seeded random states: rough field, negative-depth cells, topography, 2-D f
        rng = np.random.default_rng(1000 * M + N)
        h = 8000.0 + 500.0 * rng.uniform(-1.0, 1.0, shape)
        u = 50.0 * rng.standard_normal(shape)
        v = 50.0 * rng.standard_normal(shape)
        f = 1.0e-4 * rng.uniform(-1.0, 1.0, shape)
        hs = 300.0 * rng.uniform(0.0, 1.0, shape)
        hneg = h.copy()
        hneg.flat[rng.choice(h.size, size=max(2, h.size // 10), replace=False)] *= -1.0
        for name, (hh, tt) in {
            "rand": (h, np.zeros(shape)),
            "randhs": (h, hs),
            "randneg": (hneg, hs),
        }.items():
            for diff in (False, True):
                tag = f"{M}x{N}/{name}/d{int(diff)}"
                with np.errstate(all="ignore"):
                    hn, un, vn = one_step(ref, ll, cg, 40.0, f, tt, hh, u, v, diff)
                out[f"{tag}/h"], out[f"{tag}/u"], out[f"{tag}/v"] = hh, u, v
                out[f"{tag}/f"], out[f"{tag}/hs"] = f, tt
                out[f"{tag}/hnew"], out[f"{tag}/unew"], out[f"{tag}/vnew"] = hn, un, vn
                cases.append(tag)
    out["cases"] = np.asarray(cases)
    return out


def multi_step_fixture(ref):
    """Full driver runs with every step saved (save interval 1)."""
    out = {}
    runs = [
        # M, N, ic, sim_days, courant, diffusion
        (10, 10, 0, 0.05, 0.5, True),
        (10, 10, 1, 0.05, 0.5, False),
        (16, 12, 0, 0.03, 0.5, False),
        (16, 12, 1, 0.03, 0.4, True),
        (24, 17, 0, 0.02, 0.5, True),  # odd N -> N+1 (grid.py:83)
    ]
    meta = []
    for k, (M, N, ic, T, cfl, diff) in enumerate(runs):
        sd = {"interval": 1}
        h, u, v = ref["numpy"].solve(M, N, ref["ic"].ICType(ic), T, cfl, diff, save_data=sd)
        for name in ("h", "u", "v", "t", "phi", "theta"):
            out[f"run{k}/{name}"] = sd[name]
        out[f"run{k}/h_final"], out[f"run{k}/u_final"], out[f"run{k}/v_final"] = h, u, v
        meta.append((M, N, ic, T, cfl, int(diff)))
    out["runs"] = np.asarray(meta, dtype=np.float64)
    return out


def main():
    ref = load_reference()
    data = {}
    for var in ("h", "u", "v", "t", "phi", "theta"):
        data[var] = np.load(os.path.join(REF, "data", f"swes-numpy-ref-10-{var}.npy"))
    np.savez_compressed(os.path.join(HERE, "ref_data_10.npz"), **data)
    np.savez_compressed(os.path.join(HERE, "grid_ic.npz"), **grid_ic_fixture(ref))
    np.savez_compressed(os.path.join(HERE, "single_step.npz"), **single_step_fixture(ref))
    np.savez_compressed(os.path.join(HERE, "multi_step.npz"), **multi_step_fixture(ref))
    # the live reference re-run of the golden configuration (numpy version recorded)
    sd = {"interval": 392}
    ref["numpy"].solve(10, 10, ref["ic"].ICType.RossbyHaurwitzWave, 2, 0.5, True, save_data=sd)
    np.savez_compressed(
        os.path.join(HERE, "ref_live_10.npz"),
        numpy_version=np.asarray(np.__version__),
        **{k: sd[k] for k in ("h", "u", "v", "t", "phi", "theta")},
    )
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
