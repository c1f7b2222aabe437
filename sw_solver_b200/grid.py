"""Lat-lon channel grid and its Cartesian metrics.

Same public names and array shapes as the reference's ``src/sw_solver/grid.py``
(``LatLonGrid`` 56-113, ``CartesianGrid`` 11-53), but built B200-first: every metric
term of this grid is separable, so the grid keeps only 1-D vectors (what the CUDA
kernels consume, see ``LatLonGrid.tables``) and exposes the reference's 2-D attributes
as on-demand views.  Small grids hand out real arrays; large ones hand out zero-copy
broadcast views, so a 16384 x 8192 grid costs kilobytes instead of the ~16 GB the
reference layout would need.  All values are bit-identical to the reference's
(tests/test_host.py checks that against reference-generated fixtures).
"""

import math
from typing import Dict, Tuple

import numpy as np

from .utils import EARTH_CONSTANTS

#: grids with at most this many points materialise real (writable, contiguous) 2-D arrays
_MATERIALISE_LIMIT = 1 << 22
#: longitude rows handled at a time when a reduction over the 2-D dx field is needed
_ROW_BLOCK = 256


def _expand(vec: np.ndarray, shape: Tuple[int, int], axis: int, real: bool) -> np.ndarray:
    """Broadcast a 1-D vector living on ``axis`` to ``shape``."""
    v = vec[:, None] if axis == 0 else vec[None, :]
    out = np.broadcast_to(v, shape)
    return np.array(out) if real else out


class LatLonGrid:
    """Latitude-longitude grid of the 85S-85N channel (reference grid.py:56-113).

    ``M`` longitudes (plus three redundant periodic columns), ``N`` latitudes (made even).
    2-D attributes: ``phi, theta, c`` of shape (M+3, N); ``cMidy, tgMidy`` (M+1, N-1);
    ``tg`` (M+1, N-2); ``tgMidx`` (M+2, N-2).
    """

    def __init__(self, M: int, N: int):
        if not (M > 1 and N > 1):  # grid.py:76-80
            raise ValueError(
                f"Number of grid points along each direction must be greater than one. Found {(M, N)}"
            )
        N = N + (N % 2)  # grid.py:83: even number of latitudes
        self.M, self.N = int(M), int(N)
        dphi = 2.0 * math.pi / M
        self.dphi = dphi
        # grid.py:86-87 / 93-96
        self.phi1d = np.linspace(-dphi, 2.0 * math.pi + dphi, M + 3)
        lim = 85.0 / 180.0 * math.pi
        self.theta1d = np.linspace(-lim, lim, N)
        th = self.theta1d
        thmid = 0.5 * (th[1:] + th[:-1])  # grid.py:103 operand order
        self.c1d = np.cos(th)
        self.cMidy1d = np.cos(thmid)
        self.tg1d = np.tan(th)  # grid.py:106 takes rows 1..N-2 of this
        # grid.py:107: 0.5*(theta+theta) == theta exactly, so tgMidx is tg
        self.tgMidy1d = np.tan(0.5 * (th[:-1] + th[1:]))  # grid.py:108 operand order
        self._real = (M + 3) * N <= _MATERIALISE_LIMIT
        self._cache: Dict[str, np.ndarray] = {}

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.M + 3, self.N)

    def _view(self, name, vec, shape, axis):
        if name not in self._cache:
            self._cache[name] = _expand(vec, shape, axis, self._real)
        return self._cache[name]

    @property
    def phi(self):
        return self._view("phi", self.phi1d, self.shape, 0)

    @property
    def theta(self):
        return self._view("theta", self.theta1d, self.shape, 1)

    @property
    def c(self):
        return self._view("c", self.c1d, self.shape, 1)

    @property
    def cMidy(self):
        return self._view("cMidy", self.cMidy1d, (self.M + 1, self.N - 1), 1)

    @property
    def tg(self):
        return self._view("tg", self.tg1d[1:-1], (self.M + 1, self.N - 2), 1)

    @property
    def tgMidx(self):
        return self._view("tgMidx", self.tg1d[1:-1], (self.M + 2, self.N - 2), 1)

    @property
    def tgMidy(self):
        return self._view("tgMidy", self.tgMidy1d, (self.M + 1, self.N - 1), 1)


class CartesianGrid:
    """Cartesian metrics of a ``LatLonGrid`` (reference grid.py:11-53).

    ``x = (a cos(theta)) * phi`` is the only genuinely 2-D quantity (its roundoff noise is
    part of the numerical contract: ``dx`` is a difference of rounded products), so it is
    described by the two vectors ``ac1d`` and ``phi1d``; ``y``, ``dy``, ``dy1`` and the
    centred y-increments depend on latitude only.
    """

    def __init__(self, latlon_grid: LatLonGrid):
        g = latlon_grid
        self.M, self.N = g.M, g.N
        self._real = g._real
        self._cache: Dict[str, np.ndarray] = {}
        a = EARTH_CONSTANTS.a
        self.phi1d = g.phi1d
        self.ac1d = a * g.c1d  # grid.py:35: (a*cos(theta))*phi
        self.y1d = a * g.theta1d  # grid.py:36
        y1 = a * np.sin(g.theta1d)  # grid.py:37
        self.dy1d = self.y1d[1:] - self.y1d[:-1]  # grid.py:41
        self.dy11d = y1[1:] - y1[:-1]  # grid.py:42
        self.dyc1d = 0.5 * (self.dy1d[:-1] + self.dy1d[1:])  # grid.py:52 (rows 1..N-2)
        self.dy1c1d = 0.5 * (self.dy11d[:-1] + self.dy11d[1:])  # grid.py:53
        self.dymin = self.dy1d.min()  # grid.py:47
        self.dxmin = self._dx_min()  # grid.py:46

    # --- the 2-D x family, evaluated blockwise ---------------------------------------
    def _x_rows(self, i0: int, i1: int) -> np.ndarray:
        return self.phi1d[i0:i1, None] * self.ac1d[None, :]

    def _dx_min(self):
        """min over the whole 2-D dx field (grid.py:46) without forming it.

        dx[i, j] = ac_j*phi_{i+1} - ac_j*phi_i equals ac_j*dphi up to a relative roundoff of
        ~1e-10 at most, whereas ac_j changes from one latitude to the next by a relative
        ~34/(N-1) next to the channel walls: the minimum therefore sits in one of the two
        outermost latitudes on either side, and scanning those four columns over all
        longitudes returns the reference's value bit for bit (checked against the full
        scan in tests/test_host.py).
        """
        cols = sorted({0, 1, self.N - 2, self.N - 1})
        x = self.phi1d[:, None] * self.ac1d[None, cols]
        return np.float64((x[1:] - x[:-1]).min())

    def _dx_min_full(self):
        """The literal blockwise scan of all (M+2) x N increments (test reference)."""
        best = np.inf
        nrow = self.M + 3
        for i0 in range(0, nrow - 1, _ROW_BLOCK):
            i1 = min(nrow, i0 + _ROW_BLOCK + 1)
            x = self._x_rows(i0, i1)
            best = min(best, (x[1:] - x[:-1]).min())
        return np.float64(best)

    def _get(self, name, build):
        if name in self._cache:
            return self._cache[name]
        val = build()
        if self._real:
            self._cache[name] = val
        return val

    @property
    def x(self):
        return self._get("x", lambda: self._x_rows(0, self.M + 3))

    @property
    def dx(self):
        def build():
            x = self.x
            return x[1:, :] - x[:-1, :]  # grid.py:40

        return self._get("dx", build)

    @property
    def dxc(self):
        def build():
            dx = self.dx
            return 0.5 * (dx[:-1, 1:-1] + dx[1:, 1:-1])  # grid.py:51

        return self._get("dxc", build)

    # --- latitude-only family as 2-D views --------------------------------------------
    def _view(self, name, vec, shape):
        if name not in self._cache:
            self._cache[name] = _expand(vec, shape, 1, self._real)
        return self._cache[name]

    @property
    def y(self):
        return self._view("y", self.y1d, (self.M + 3, self.N))

    @property
    def dy(self):
        return self._view("dy", self.dy1d, (self.M + 3, self.N - 1))

    @property
    def dy1(self):
        return self._view("dy1", self.dy11d, (self.M + 3, self.N - 1))

    @property
    def dyc(self):
        return self._view("dyc", self.dyc1d, (self.M + 1, self.N - 2))

    @property
    def dy1c(self):
        return self._view("dy1c", self.dy1c1d, (self.M + 1, self.N - 2))
