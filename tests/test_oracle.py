"""Pins the CPU oracle (oracle/) to the reference: bit-for-bit against fixtures generated
by importing the reference numpy solver (tests/golden/make_golden.py) and against the
reference's own known-answer files (data/swes-numpy-ref-10-*.npy)."""

import numpy as np
import pytest

from oracle import sw_oracle as so
from conftest import normwise


def _bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_grid_and_ic_bit_exact(golden_grid_ic):
    g = golden_grid_ic
    for M, N in g["sizes"]:
        tag = f"{M}x{N}"
        og = so.OracleGrid(int(M), int(N))
        for name in ("phi", "theta", "c", "cMidy", "tg", "tgMidx", "tgMidy"):
            assert _bits(getattr(og, name), g[f"{tag}/ll/{name}"]), (tag, name)
        for name in ("x", "y", "dx", "dy", "dy1", "dxc", "dyc", "dy1c"):
            assert _bits(getattr(og, name), g[f"{tag}/cg/{name}"]), (tag, name)
        assert og.dxmin == g[f"{tag}/cg/dxmin"] and og.dymin == g[f"{tag}/cg/dymin"]
        for name, arr in og.diffusion_coefficients().items():
            assert _bits(arr, g[f"{tag}/dc/{name}"]), (tag, name)
        for ic in (0, 1):
            for name, arr in zip("huvf", so.initial_conditions(ic, og)):
                assert _bits(arr, g[f"{tag}/ic{ic}/{name}"]), (tag, ic, name)


def test_single_step_bit_exact(golden_single_step, golden_grid_ic):
    g = golden_single_step
    for tag in g["cases"]:
        tag = str(tag)
        size, kind, d = tag.split("/")
        M, N = (int(x) for x in size.split("x"))
        og = so.OracleGrid(M, N)
        diff = d == "d1"
        if kind.startswith("ic"):
            h, u, v, f = so.initial_conditions(int(kind[2:]), og)
            hs, dt = np.zeros(og.shape), 1.25
        else:
            h, u, v, f, hs = (g[f"{tag}/{k}"] for k in ("h", "u", "v", "f", "hs"))
            dt = 40.0
        before = [x.copy() for x in (h, u, v, f, hs)]
        hn, un, vn = so.lax_wendroff_update(og, dt, f, hs, h, u, v, diff)
        for name, arr in zip(("hnew", "unew", "vnew"), (hn, un, vn)):
            assert _bits(arr, g[f"{tag}/{name}"]), (tag, name)
        # purity, reference tests/numpy_test.py:33-80
        for x0, x1 in zip(before, (h, u, v, f, hs)):
            assert _bits(x0, x1)


def test_multi_step_bit_exact(golden_multi_step):
    g = golden_multi_step
    for k, (M, N, ic, T, cfl, diff) in enumerate(g["runs"]):
        sd = {"interval": 1}
        h, u, v = so.solve(int(M), int(N), int(ic), float(T), float(cfl), bool(diff), save_data=sd)
        for name in ("t", "phi", "theta", "h", "u", "v"):
            assert _bits(sd[name], g[f"run{k}/{name}"]), (k, name)
        for name, arr in zip(("h_final", "u_final", "v_final"), (h, u, v)):
            assert _bits(arr, g[f"run{k}/{name}"]), (k, name)


def test_reference_known_answer_files(golden_ref10):
    """The reference's own regression (tests/numpy_test.py:9-30): 10x10, RH, 2 days,
    CFL 0.5, diffusion on, one save at step 392 -- np.allclose as the reference asserts,
    plus the tighter normwise bound recorded in SURVEY.md section 4."""
    sd = {"interval": 392}
    so.solve(10, 10, 0, 2, 0.5, True, save_data=sd)
    assert len(sd["_dts"]) == 392
    for name in ("h", "u", "v", "t", "phi", "theta"):
        assert sd[name].shape == golden_ref10[name].shape
        assert np.allclose(sd[name], golden_ref10[name]), name
    for name in ("t", "phi", "theta"):
        assert _bits(sd[name], golden_ref10[name])
    for name in ("h", "u", "v"):
        assert normwise(sd[name], golden_ref10[name]) < 1e-12


def test_live_reference_rerun_bit_exact():
    """Same configuration against the reference re-run under today's numpy."""
    from conftest import load_golden

    live = load_golden("ref_live_10.npz")
    sd = {"interval": 392}
    so.solve(10, 10, 0, 2, 0.5, True, save_data=sd)
    for name in ("h", "u", "v", "t"):
        assert _bits(sd[name], live[name]), name


def test_driver_semantics():
    with pytest.raises(ValueError):
        so.solve(1, 10, 0, 1, 0.5)
    with pytest.raises(ValueError):
        so.solve(10, 10, 0, -1.0, 0.5)
    h, u, v = so.solve(6, 6, 0, 0.0, 0.5)
    og = so.OracleGrid(6, 6)
    h0, u0, v0, _ = so.initial_conditions(0, og)
    assert _bits(h, h0) and _bits(u, u0) and _bits(v, v0)


def test_oracle_matches_the_staged_reference_live():
    """oracle/_ref holds the UNMODIFIED reference numpy modules (staged by __graft_entry__.build()
    where /root/reference exists; it travels to the GPU box like the built shared objects).  When
    present, the C oracle must reproduce the reference's own `solve` bit for bit, live: both ICs,
    with and without diffusion, odd N, the final-time clamp."""
    from oracle import ref_numpy as rn

    if not rn.available():
        pytest.skip("oracle/_ref is not staged on this machine")
    for (M, N, ic, diff, nsteps) in ((10, 10, 0, True, 25), (12, 9, 1, False, 12), (33, 20, 0, False, 8), (16, 12, 1, True, 6)):
        sd = {}
        so.solve(M, N, ic, 1e9, 0.5, diff, save_data=sd, max_steps=nsteps)
        h, u, v, final_time, _ = rn.solve_steps(M, N, ic, nsteps, 0.5, diff, sd["_dts"])
        sd2 = {}
        ho, uo, vo = so.solve(M, N, ic, final_time / 86400.0, 0.5, diff, save_data=sd2)
        assert len(sd2["_dts"]) == nsteps
        assert _bits(h, ho) and _bits(u, uo) and _bits(v, vo), (M, N, ic, diff)
