"""Python face of the CPU oracle -- TEST INFRASTRUCTURE, not the product.

A plain restatement of the reference numpy solver's path (ai2cm/sw_solver):

* grid metrics        -> ``OracleGrid``               (src/sw_solver/grid.py:35-53, 86-108)
* initial conditions  -> ``initial_conditions``       (src/sw_solver/ic.py:46-154)
* diffusion weights   -> ``diffusion_coefficients``   (src/sw_solver/numpy.py:27-63)
* one loop body       -> ``libsw_oracle.so`` (C)      (src/sw_solver/numpy.py:503-549)
* driver              -> ``solve``                    (src/sw_solver/numpy.py:409-575)

Everything is laid out as the reference lays it out (2-D float64 arrays, state
``(M+3, N)`` with latitude contiguous).  Parity of this oracle is PINNED: tests/test_oracle.py
checks it bit-for-bit against fixtures produced by importing the reference itself
(tests/golden/make_golden.py) and against the reference's own known-answer files
(data/swes-numpy-ref-10-*.npy, stored as tests/golden/ref_data_10.npz).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsw_oracle.so")

G_GRAV = 9.80616  # utils.py:28
A_EARTH = 6.37122e6  # utils.py:30
OMEGA = 7.292e-5  # utils.py:31
NU = 5.0e5  # utils.py:33

_dp = ctypes.POINTER(ctypes.c_double)


class _CGrid(ctypes.Structure):
    _fields_ = (
        [("M", ctypes.c_int), ("N", ctypes.c_int)]
        + [(k, ctypes.c_double) for k in ("g", "a", "nu")]
        + [
            (k, _dp)
            for k in (
                "c", "cMidy", "tg", "tgMidx", "tgMidy",
                "dx", "dy", "dy1", "dxc", "dyc", "dy1c",
                "Ax", "Bx", "Cx", "Ay", "By", "Cy",
            )
        ]
    )


def build(force: bool = False) -> str:
    """Compile oracle/sw_oracle.c (gcc, -ffp-contract=off) into oracle/_build/."""
    src = os.path.join(_HERE, "sw_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.swo_lw_update.restype = None
        L.swo_lw_update.argtypes = [ctypes.POINTER(_CGrid), ctypes.c_double] + [_dp] * 5 + [ctypes.c_int] + [_dp] * 3
        L.swo_apply_bcs.restype = None
        L.swo_apply_bcs.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.swo_cfl_eigen.restype = None
        L.swo_cfl_eigen.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, _dp, _dp, _dp, _dp]
        L.swo_cfl_dt.restype = ctypes.c_double
        L.swo_cfl_dt.argtypes = [ctypes.c_double] * 5
        L.swo_step.restype = ctypes.c_double
        L.swo_step.argtypes = (
            [ctypes.POINTER(_CGrid)] + [ctypes.c_double] * 4 + [_dp] + [_dp] * 5 + [ctypes.c_int] + [_dp] * 3
        )
        L.swo_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _c64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class OracleGrid:
    """Lat-lon channel grid and its Cartesian metrics, as 2-D arrays.

    grid.py:74-108 (LatLonGrid) and grid.py:32-53 (CartesianGrid).
    """

    def __init__(self, M: int, N: int):
        if not (M > 1 and N > 1):  # grid.py:76-80
            raise ValueError(f"need more than one point per direction, got {(M, N)}")
        N += N % 2  # grid.py:83
        self.M, self.N = M, N
        dphi = 2.0 * math.pi / M
        lon = np.linspace(-dphi, 2.0 * math.pi + dphi, M + 3)  # grid.py:86-87
        lim = 85.0 / 180.0 * math.pi
        lat = np.linspace(-lim, lim, N)  # grid.py:93-96
        self.phi, self.theta = np.meshgrid(lon, lat, indexing="ij")
        th = self.theta
        self.c = np.cos(th)  # grid.py:102
        self.cMidy = np.cos(0.5 * (th[1:-1, 1:] + th[1:-1, :-1]))  # grid.py:103
        self.tg = np.tan(th[1:-1, 1:-1])  # grid.py:106
        self.tgMidx = np.tan(0.5 * (th[:-1, 1:-1] + th[1:, 1:-1]))  # grid.py:107
        self.tgMidy = np.tan(0.5 * (th[1:-1, :-1] + th[1:-1, 1:]))  # grid.py:108
        # Cartesian metrics, grid.py:35-53
        self.x = A_EARTH * np.cos(th) * self.phi
        self.y = A_EARTH * th
        y1 = A_EARTH * np.sin(th)
        self.dx = np.diff(self.x, axis=0)
        self.dy = np.diff(self.y, axis=1)
        self.dy1 = np.diff(y1, axis=1)
        self.dxmin = self.dx.min()
        self.dymin = self.dy.min()
        self.dxc = 0.5 * (self.dx[:-1, 1:-1] + self.dx[1:, 1:-1])
        self.dyc = 0.5 * (self.dy[1:-1, :-1] + self.dy[1:-1, 1:])
        self.dy1c = 0.5 * (self.dy1[1:-1, :-1] + self.dy1[1:-1, 1:])
        self._diff = None
        self._keep = []

    @property
    def shape(self):
        return self.phi.shape

    def diffusion_coefficients(self):
        if self._diff is None:
            self._diff = diffusion_coefficients(self)
        return self._diff

    def cstruct(self, with_diffusion: bool):
        g = _CGrid()
        g.M, g.N = self.M, self.N
        g.g, g.a, g.nu = G_GRAV, A_EARTH, NU
        keep = []
        for k in ("c", "cMidy", "tg", "tgMidx", "tgMidy", "dx", "dy", "dy1", "dxc", "dyc", "dy1c"):
            arr = _c64(getattr(self, k))
            keep.append(arr)
            setattr(g, k, _p(arr))
        if with_diffusion:
            for k, arr in self.diffusion_coefficients().items():
                arr = _c64(arr)
                keep.append(arr)
                setattr(g, k, _p(arr))
        self._keep.append(keep)
        return g


def _three_point_weights(d0, d1):
    """Weights (centre, +1, -1) of the 3-point first derivative on a non-uniform
    grid with spacing d0 on the minus side and d1 on the plus side
    (numpy.py:32-34, 37-39, 42-44 and the y twins 50-62)."""
    return (d1 - d0) / (d1 * d0), d0 / (d1 * (d1 + d0)), -d1 / (d0 * (d1 + d0))


def diffusion_coefficients(grid: OracleGrid):
    """numpy.py:27-63: x-weights padded periodically, y-weights by edge duplication."""
    ax, bx, cx = _three_point_weights(grid.dx[:-1, 1:-1], grid.dx[1:, 1:-1])
    ay, by, cy = _three_point_weights(grid.dy[1:-1, :-1], grid.dy[1:-1, 1:])

    def wrap_x(w):  # numpy.py:35, 40, 45
        return np.concatenate((w[-2:-1, :], w, w[1:2, :]), axis=0)

    def edge_y(w):  # numpy.py:53, 58, 63
        return np.concatenate((w[:, 0:1], w, w[:, -1:]), axis=1)

    return {
        "Ax": wrap_x(ax), "Bx": wrap_x(bx), "Cx": wrap_x(cx),
        "Ay": edge_y(ay), "By": edge_y(by), "Cy": edge_y(cy),
    }


def initial_conditions(ic_type: int, grid: OracleGrid):
    """Williamson test cases 6 (ic_type 0) and 2 (ic_type 1); ic.py:46-154."""
    th, ph = grid.theta, grid.phi
    a, om, g = A_EARTH, OMEGA, G_GRAV
    f = 2.0 * om * np.sin(th)  # ic.py:46
    ct, st = np.cos(th), np.sin(th)
    if int(ic_type) == 0:  # Rossby-Haurwitz wave, ic.py:53-107
        w = K = 7.848e-6
        h0, R = 8e3, 4.0
        A = 0.5 * w * (2.0 * om + w) * (ct ** 2.0) + 0.25 * (K ** 2.0) * (ct ** (2.0 * R)) * (
            (R + 1.0) * (ct ** 2.0) + (2.0 * (R ** 2.0) - R - 2.0) - 2.0 * (R ** 2.0) * (ct ** (-2.0))
        )
        B = (2.0 * (om + w) * K) / ((R + 1.0) * (R + 2.0)) * (ct ** R) * (
            ((R ** 2.0) + 2.0 * R + 2.0) - ((R + 1.0) ** 2.0) * (ct ** 2.0)
        )
        C = 0.25 * (K ** 2.0) * (ct ** (2.0 * R)) * ((R + 1.0) * (ct ** 2.0) - (R + 2.0))
        a2 = a ** 2.0
        h = h0 + (a2 * A + a2 * B * np.cos(R * ph) + a2 * C * np.cos(2.0 * R * ph)) / g
        u = a * w * ct + a * K * (ct ** (R - 1.0)) * (R * (st ** 2.0) - (ct ** 2.0)) * np.cos(R * ph)
        v = -a * K * R * (ct ** (R - 1.0)) * st * np.sin(R * ph)
    elif int(ic_type) == 1:  # zonal geostrophic flow, alpha = pi/2, ic.py:119-154
        alpha = math.pi / 2
        u0 = 2.0 * math.pi * a / (12.0 * 24.0 * 3600.0)
        h0 = 2.94e4 / g
        rot = -np.cos(ph) * ct * np.sin(alpha) + st * np.cos(alpha)
        f = 2.0 * om * rot
        h = h0 - (a * om * u0 + 0.5 * (u0 ** 2.0)) * (rot ** 2.0) / g
        u = u0 * (ct * np.cos(alpha) + np.cos(ph) * st * np.sin(alpha))
        v = -u0 * np.sin(ph) * np.sin(alpha)
    else:
        raise ValueError(f"unknown ic_type {ic_type}")
    return h, u, v, f


def lax_wendroff_update(grid: OracleGrid, dt, f, hs, h, u, v, use_diffusion=False):
    """One LW update (+ optional diffusion increment): numpy.py:136-356, 541-544.

    Returns new (M+1, N-2) arrays; inputs untouched.
    """
    M, N = grid.M, grid.N
    f, hs, h, u, v = (_c64(x) for x in (f, hs, h, u, v))
    out = [np.empty((M + 1, N - 2)) for _ in range(3)]
    cg = grid.cstruct(use_diffusion)
    lib().swo_lw_update(ctypes.byref(cg), float(dt), _p(f), _p(hs), _p(h), _p(u), _p(v),
                        int(bool(use_diffusion)), *(_p(o) for o in out))
    return tuple(out)


def apply_bcs(q, qnew):
    """numpy.py:385-406 (q modified in place and returned)."""
    assert q.flags.c_contiguous and q.dtype == np.float64
    qnew = _c64(qnew)
    lib().swo_apply_bcs(q.shape[0] - 3, q.shape[1], _p(q), _p(qnew))
    return q


def cfl_eigen(h, u, v):
    """numpy.py:503-521 over the full ghosted arrays."""
    h, u, v = (_c64(x) for x in (h, u, v))
    ex, ey = ctypes.c_double(), ctypes.c_double()
    lib().swo_cfl_eigen(h.shape[0] - 3, h.shape[1], G_GRAV, _p(h), _p(u), _p(v),
                        ctypes.byref(ex), ctypes.byref(ey))
    return ex.value, ey.value


class Stepper:
    """Re-usable loop body (numpy.py:503-549) on a fixed grid; state updated in place."""

    def __init__(self, grid: OracleGrid, f, hs=None, use_diffusion=True):
        self.grid = grid
        self.f = _c64(f)
        self.hs = _c64(hs if hs is not None else np.zeros(grid.shape))
        self.use_diffusion = bool(use_diffusion)
        self.cg = grid.cstruct(self.use_diffusion)
        self.work = [np.empty((grid.M + 1, grid.N - 2)) for _ in range(3)]

    def step(self, h, u, v, courant, time, final_time):
        t = ctypes.c_double(time)
        dt = lib().swo_step(ctypes.byref(self.cg), float(courant), float(self.grid.dxmin),
                            float(self.grid.dymin), float(final_time), ctypes.byref(t),
                            _p(self.f), _p(self.hs), _p(h), _p(u), _p(v),
                            int(self.use_diffusion), *(_p(w) for w in self.work))
        return dt, t.value


def solve(num_lon_pts, num_lat_pts, ic_type, sim_days, courant, use_diffusion=True,
          save_data=None, print_interval=-1, max_steps=None):
    """Driver with the reference's semantics, numpy.py:409-575.

    ``max_steps`` (oracle-only extension) stops after that many steps.
    """
    save_data = save_data if save_data is not None else {}
    grid = OracleGrid(num_lon_pts, num_lat_pts)
    if sim_days < 0.0:
        raise ValueError(f"Final time {sim_days} must be non-negative.")
    final_time = 24 * 3600 * sim_days
    h, u, v, f = initial_conditions(ic_type, grid)
    h, u, v = (_c64(x) for x in (h, u, v))
    stepper = Stepper(grid, f, None, use_diffusion)
    interval = save_data.get("interval", 0)
    if interval > 0:
        ts = [0.0]
        snaps = [[q[1:-1, :].copy()] for q in (h, u, v)]
    n, time, dts = 0, 0.0, []
    while time < final_time and (max_steps is None or n < max_steps):
        n += 1
        dt, time = stepper.step(h, u, v, courant, time, final_time)
        dts.append(dt)
        if print_interval > 0 and n % print_interval == 0:
            print(f"Time = {time / 3600.0:6.2f} hours (max {int(final_time / 3600.0)}); "
                  f"max(|u|) = {np.sqrt(u * u + v * v).max():16.16f}")
        if interval > 0 and n % interval == 0:
            ts.append(time)
            for s, q in zip(snaps, (h, u, v)):
                s.append(q[1:-1, :].copy())
    if interval > 0:
        for k, s in zip("huv", snaps):
            save_data[k] = np.stack(s, axis=2)
        save_data["t"] = np.asarray(ts)
        save_data["phi"], save_data["theta"] = grid.phi, grid.theta
    save_data["_dts"] = np.asarray(dts)
    return h, u, v


def num_threads() -> int:
    return lib().swo_num_threads()
