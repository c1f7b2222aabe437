"""CPU-only tests of the host layer: grid / IC utilities against reference-generated
fixtures (bit-for-bit), band decomposition, latitude tables, and that libswb200.so loads
and exports every symbol include/swb200.h declares (no compute calls: there is no GPU here)."""

import ctypes
import os
import re

import numpy as np
import pytest

import sw_solver_b200 as swb
from sw_solver_b200 import _lib
from sw_solver_b200.grid import CartesianGrid, LatLonGrid
from sw_solver_b200.ic import ICType, get_initial_conditions, separable_rossby_haurwitz
from oracle import sw_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_grid_bit_exact_with_reference(golden_grid_ic):
    g = golden_grid_ic
    for M, N in g["sizes"]:
        tag = f"{M}x{N}"
        ll = LatLonGrid(int(M), int(N))
        cg = CartesianGrid(ll)
        for name in ("phi", "theta", "c", "cMidy", "tg", "tgMidx", "tgMidy"):
            assert _bits(getattr(ll, name), g[f"{tag}/ll/{name}"]), (tag, name)
        for name in ("x", "y", "dx", "dy", "dy1", "dxc", "dyc", "dy1c"):
            assert _bits(getattr(cg, name), g[f"{tag}/cg/{name}"]), (tag, name)
        assert cg.dxmin == g[f"{tag}/cg/dxmin"] and cg.dymin == g[f"{tag}/cg/dymin"]
        assert ll.shape == tuple(g[f"{tag}/ll/phi"].shape)


def test_ic_bit_exact_with_reference(golden_grid_ic):
    g = golden_grid_ic
    for M, N in g["sizes"]:
        tag = f"{M}x{N}"
        ll = LatLonGrid(int(M), int(N))
        for ic in ICType:
            out = get_initial_conditions(ic, ll)
            for name, arr in zip("huvf", out):
                assert _bits(arr, g[f"{tag}/ic{int(ic)}/{name}"]), (tag, ic, name)
            for arr in out[:3]:
                assert arr.flags.c_contiguous and arr.flags.writeable and arr.dtype == np.float64


def test_separable_forms_reproduce_the_reference_zonal_flow(golden_grid_ic):
    """Williamson case 2 from 1-D tables: the separable Coriolis parameter and the separable state factors (what the
    kernels and the device-side initial state are fed) give the reference's arrays bit for bit (ic.py:119-154)."""
    from sw_solver_b200.ic import SeparableCoriolis, separable_zonal_flow, zonal_flow_coriolis

    g = golden_grid_ic
    for M, N in g["sizes"]:
        tag = f"{M}x{N}"
        ll = LatLonGrid(int(M), int(N))
        ref = {k: g[f"{tag}/ic1/{k}"] for k in "huvf"}
        sc = zonal_flow_coriolis(ll)
        assert isinstance(sc, SeparableCoriolis) and sc.matches(ref["f"]) and _bits(sc.array(), ref["f"])
        assert not sc.matches(ref["f"] + 1e-20 * np.abs(ref["f"]).max()) or np.array_equal(ref["f"] + 1e-20 * np.abs(ref["f"]).max(), ref["f"])
        assert not sc.matches(ref["f"][:, :-1])
        s = separable_zonal_flow(ll)
        rot = (s["ncp"][:, None] * s["ct"][None, :]) * s["sa"] + s["sca"][None, :]
        assert _bits(s["h0"] - (s["K"] * (rot * rot)) / s["g"], ref["h"])
        assert _bits(s["u0"] * (s["cca"][None, :] + (s["cp"][:, None] * s["st"][None, :]) * s["sa"]), ref["u"])
        assert _bits(np.broadcast_to((s["sp_nu0"] * s["sa"])[:, None], ref["v"].shape), ref["v"])
    # a field that is not the zonal flow's is not mistaken for it
    ll = LatLonGrid(12, 8)
    f = get_initial_conditions(ICType.ZonalGeostrophicFlow, ll)[3]
    f2 = f.copy()
    f2[3, 4] = np.nextafter(f2[3, 4], np.inf)
    assert zonal_flow_coriolis(ll).matches(f) and not zonal_flow_coriolis(ll).matches(f2)


def test_lazy_views_on_large_grids_match_small_grid_semantics():
    """Above the materialisation limit 2-D attributes are zero-copy broadcast views with the
    same values the oracle's dense arrays hold."""
    M, N = 3000, 1500  # (M+3)*N > 2**22
    ll = LatLonGrid(M, N)
    cg = CartesianGrid(ll)
    assert ll.phi.strides[1] == 0 and ll.theta.strides[0] == 0
    og = so.OracleGrid(M, N)
    assert _bits(ll.phi, og.phi) and _bits(ll.theta, og.theta) and _bits(ll.cMidy, og.cMidy)
    assert _bits(cg.dy1, og.dy1) and _bits(cg.dyc, og.dyc)
    assert cg.dxmin == og.dxmin and cg.dymin == og.dymin
    # a strip of the genuinely 2-D family
    assert _bits(cg._x_rows(100, 110), og.x[100:110])
    s = separable_rossby_haurwitz(ll)
    h, u, v, _ = so.initial_conditions(0, og)
    i = np.array([0, 1, 777, M + 2])[:, None]
    assert _bits(s["h0"] + (s["hA"][None, :] + s["hB"][None, :] * s["cosR"][i] + s["hC"][None, :] * s["cos2R"][i]) / s["g"], h[i[:, 0]])
    assert _bits(s["uA"][None, :] + s["uB"][None, :] * s["cosR"][i], u[i[:, 0]])
    assert _bits(s["vB"][None, :] * s["sinR"][i], v[i[:, 0]])


def test_grid_errors():
    with pytest.raises(ValueError):
        LatLonGrid(1, 10)
    with pytest.raises(ValueError):
        LatLonGrid(10, 1)
    assert LatLonGrid(7, 9).N == 10  # odd N -> N + 1 (grid.py:83)


def test_diffusion_coefficients_match_reference(golden_grid_ic):
    from sw_solver_b200.solver import DiffusionCoefficients

    g = golden_grid_ic
    for M, N in g["sizes"]:
        tag = f"{M}x{N}"
        dc = DiffusionCoefficients(CartesianGrid(LatLonGrid(int(M), int(N))))
        for name in ("Ax", "Bx", "Cx", "Ay", "By", "Cy"):
            assert _bits(getattr(dc, name), g[f"{tag}/dc/{name}"]), (tag, name)


def test_band_layout_covers_interior_once():
    from sw_solver_b200.solver import band_layout

    for N, world, halo in [(10, 1, 1), (10, 2, 1), (8192, 8, 1), (8192, 8, 2), (722, 4, 2), (20, 3, 2)]:
        bands = band_layout(N, world, halo)
        rows = []
        for b in bands:
            rows += list(range(b["j_first"], b["j_last"] + 1))
            assert b["j_begin"] <= b["j_first"] - 1 and b["j_begin"] + b["n_local"] >= b["j_last"] + 2
            assert b["j_begin"] >= 0 and b["j_begin"] + b["n_local"] <= N
        assert rows == list(range(1, N - 1))
        assert bands[0]["j_begin"] == 0 and bands[-1]["j_begin"] + bands[-1]["n_local"] == N
        sizes = [b["j_last"] - b["j_first"] + 1 for b in bands]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        band_layout(6, 8, 1)


def test_latitude_tables_slice_consistently():
    from sw_solver_b200.solver import band_layout, latitude_tables
    from sw_solver_b200.ic import coriolis_latitude

    ll = LatLonGrid(16, 24)
    cg = CartesianGrid(ll)
    full = latitude_tables(ll, cg, coriolis_latitude(ll), True)
    og = so.OracleGrid(16, 24)
    assert _bits(full["c"], og.c[0]) and _bits(full["tg"][1:-1], og.tg[0]) and _bits(full["ac"] * ll.phi1d[3], og.x[3])
    assert _bits(full["tgMidy"][:-1], og.tgMidy[0]) and _bits(full["cMidy"][:-1], og.cMidy[0])
    assert _bits(full["dy"][:-1], og.dy[0]) and _bits(full["dy1"][:-1], og.dy1[0])
    assert _bits(full["dyc"][1:-1], og.dyc[0]) and _bits(full["dy1c"][1:-1], og.dy1c[0])
    d = og.diffusion_coefficients()
    assert _bits(full["Ay"], d["Ay"][0]) and _bits(full["By"], d["By"][0]) and _bits(full["Cy"], d["Cy"][0])
    for b in band_layout(24, 3, 2):
        part = latitude_tables(ll, cg, coriolis_latitude(ll), True, b["j_begin"], b["n_local"])
        for k, v in part.items():
            if k != "phi":
                assert _bits(v, full[k][b["j_begin"]:b["j_begin"] + b["n_local"]]), k


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "swb200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*\*?(swb_\w+)\(", header, flags=re.M))
    assert declared, "no prototypes parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    L = _lib.load()  # dlopen + prototype declaration (raises if a symbol is missing)
    for name in declared:
        assert hasattr(L, name)
    assert L.swb_abi_version() == 3
    assert ctypes.sizeof(_lib.SwbStatus) == 56 and ctypes.sizeof(_lib.SwbConfig) == 20 * 4 + 10 * 8


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        swb.solver.solve(10, 10, ICType.RossbyHaurwitzWave, 1, 0.5)
    # argument errors keep the reference's order and types even without a device
    with pytest.raises(ValueError):
        swb.solver.solve(1, 10, ICType.RossbyHaurwitzWave, 1, 0.5)
    with pytest.raises(ValueError):
        swb.solver.solve(10, 10, ICType.RossbyHaurwitzWave, -1.0, 0.5)
    with pytest.raises(TypeError):
        swb.solver.solve(10, 10, 0, 1, 0.5)
    # the library itself refuses to create a context without a device
    L = _lib.load()
    cfg = _lib.SwbConfig()
    cfg.struct_bytes = ctypes.sizeof(cfg)
    cfg.M, cfg.N, cfg.n_local, cfg.pitch, cfg.j_first, cfg.j_last, cfg.world = 10, 10, 10, L.swb_required_pitch(10, 0), 1, 8, 1
    ctx = ctypes.c_void_p()
    assert L.swb_create(ctypes.byref(ctx), ctypes.byref(cfg)) == -3
    assert b"no CUDA device" in L.swb_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sw_solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower(), f"{fn} mentions the oracle: the product must not depend on it"
