"""Initial conditions (Williamson test cases 6 and 2) on a ``LatLonGrid``.

Same entry points as the reference's ``src/sw_solver/ic.py`` (``ICType`` 12-19,
``get_initial_conditions`` 22-159) and bit-identical values, but evaluated from the
grid's 1-D vectors: all transcendental work happens on O(M + N) points and only the final
products / sums are broadcast to (M+3, N).  ``separable_rossby_haurwitz`` exposes that
factorisation so that large grids can be assembled directly on the GPU.
"""

import enum
import math
from typing import Dict

import numpy as np

from .grid import LatLonGrid
from .utils import EARTH_CONSTANTS


class ICType(enum.IntEnum):
    """Initial condition types (reference ic.py:12-19)."""

    RossbyHaurwitzWave = 0
    """Test case 6 by Williamson."""

    ZonalGeostrophicFlow = 1
    """Test case 2 by Williamson."""


def coriolis_latitude(latlon_grid: LatLonGrid) -> np.ndarray:
    """f = 2 omega sin(theta) as a latitude vector (reference ic.py:46)."""
    return 2.0 * EARTH_CONSTANTS.omega * np.sin(latlon_grid.theta1d)


def separable_rossby_haurwitz(latlon_grid: LatLonGrid) -> Dict[str, np.ndarray]:
    """Latitude / longitude factors of the Rossby-Haurwitz wave (reference ic.py:53-107).

    With them (all products formed left to right exactly as the reference forms them):

        h[i, j] = h0 + ((hA[j] + hB[j] * cosR[i]) + hC[j] * cos2R[i]) / g
        u[i, j] = uA[j] + uB[j] * cosR[i]
        v[i, j] = vB[j] * sinR[i]
    """
    a, om, g = EARTH_CONSTANTS.a, EARTH_CONSTANTS.omega, EARTH_CONSTANTS.g
    th, ph = latlon_grid.theta1d, latlon_grid.phi1d
    w = 7.848e-6
    K = 7.848e-6
    h0 = 8e3
    R = 4.0
    ct, st = np.cos(th), np.sin(th)
    A = 0.5 * w * (2.0 * om + w) * (ct ** 2.0) + 0.25 * (K ** 2.0) * (ct ** (2.0 * R)) * (
        (R + 1.0) * (ct ** 2.0) + (2.0 * (R ** 2.0) - R - 2.0) - 2.0 * (R ** 2.0) * (ct ** (-2.0))
    )
    B = (
        (2.0 * (om + w) * K)
        / ((R + 1.0) * (R + 2.0))
        * (ct ** R)
        * (((R ** 2.0) + 2.0 * R + 2.0) - ((R + 1.0) ** 2.0) * (ct ** 2.0))
    )
    C = 0.25 * (K ** 2.0) * (ct ** (2.0 * R)) * ((R + 1.0) * (ct ** 2.0) - (R + 2.0))
    a2 = a ** 2.0
    return {
        "h0": np.float64(h0),
        "g": np.float64(g),
        "hA": a2 * A,
        "hB": a2 * B,
        "hC": a2 * C,
        "cosR": np.cos(R * ph),
        "cos2R": np.cos(2.0 * R * ph),
        "sinR": np.sin(R * ph),
        "uA": a * w * ct,
        "uB": a * K * (ct ** (R - 1.0)) * (R * (st ** 2.0) - (ct ** 2.0)),
        "vB": -a * K * R * (ct ** (R - 1.0)) * st,
    }


class SeparableCoriolis:
    """A Coriolis parameter of the form ``f[i, j] = scale * ((lon[i] * lat_b[j]) * sa + lat_d[j])``, evaluated in
    exactly this order (every operation rounded): what the reference's Williamson case 2 builds as an (M+3, N)
    array (ic.py:126-133: ``2 Omega (-cos(phi) cos(theta) sin(alpha) + sin(theta) cos(alpha))``).  The kernels
    rebuild it per cell from the three 1-D tables, so such a state needs no 2-D Coriolis array on the device and
    runs through the TMA-ring and resident kernels like a latitude-only one."""

    def __init__(self, lon: np.ndarray, lat_b: np.ndarray, lat_d: np.ndarray, scale: float, sa: float):
        self.lon = np.ascontiguousarray(lon, dtype=np.float64)
        self.lat_b = np.ascontiguousarray(lat_b, dtype=np.float64)
        self.lat_d = np.ascontiguousarray(lat_d, dtype=np.float64)
        self.scale, self.sa = float(scale), float(sa)

    def array(self) -> np.ndarray:
        """The (M+3, N) array, formed as the reference forms it."""
        return self.scale * ((self.lon[:, None] * self.lat_b[None, :]) * self.sa + self.lat_d[None, :])

    def matches(self, f) -> bool:
        f = np.asarray(f)
        return f.shape == (len(self.lon), len(self.lat_b)) and bool(np.array_equal(f, self.array()))


_ZGF_ALPHA = math.pi / 2  # ic.py:118


def zonal_flow_coriolis(latlon_grid: LatLonGrid) -> SeparableCoriolis:
    """The Coriolis parameter of the zonal geostrophic flow (reference ic.py:126-133) in separable form."""
    th, ph = latlon_grid.theta1d, latlon_grid.phi1d
    return SeparableCoriolis(-np.cos(ph), np.cos(th), np.sin(th) * np.cos(_ZGF_ALPHA), 2.0 * EARTH_CONSTANTS.omega, np.sin(_ZGF_ALPHA))


def separable_zonal_flow(latlon_grid: LatLonGrid) -> Dict[str, np.ndarray]:
    """Longitude / latitude factors of the zonal geostrophic flow (reference ic.py:119-154).  With
    ``rot[i, j] = (ncp[i] * ct[j]) * sa + sca[j]`` (the rotated sine of latitude):

        h[i, j] = h0 - (K * (rot * rot)) / g
        u[i, j] = u0 * (cca[j] + (cp[i] * st[j]) * sa)
        v[i, j] = (nu0 * sp[i]) * sa
    """
    a, om, g = EARTH_CONSTANTS.a, EARTH_CONSTANTS.omega, EARTH_CONSTANTS.g
    alpha = _ZGF_ALPHA
    u0 = 2.0 * math.pi * a / (12.0 * 24.0 * 3600.0)
    th, ph = latlon_grid.theta1d, latlon_grid.phi1d
    return {
        "h0": np.float64(2.94e4 / g), "g": np.float64(g), "K": np.float64(a * om * u0 + 0.5 * (u0 ** 2.0)),
        "u0": np.float64(u0), "sa": np.float64(np.sin(alpha)),
        "ncp": -np.cos(ph), "cp": np.cos(ph), "sp_nu0": -u0 * np.sin(ph),
        "ct": np.cos(th), "st": np.sin(th), "sca": np.sin(th) * np.cos(alpha), "cca": np.cos(th) * np.cos(alpha),
    }


def rossby_haurwitz_window(latlon_grid: LatLonGrid, j_begin: int, n_local: int):
    """h, u, v of the Rossby-Haurwitz wave on latitudes ``j_begin .. j_begin+n_local-1`` only,
    shape (M+3, n_local): bit-identical to the same columns of ``get_initial_conditions``
    (same products and sums in the same order), without forming the full grid."""
    s = separable_rossby_haurwitz(latlon_grid)
    sl = slice(j_begin, j_begin + n_local)
    cR, c2R, sR = s["cosR"][:, None], s["cos2R"][:, None], s["sinR"][:, None]
    h = s["h0"] + (s["hA"][None, sl] + s["hB"][None, sl] * cR + s["hC"][None, sl] * c2R) / s["g"]
    u = s["uA"][None, sl] + s["uB"][None, sl] * cR
    v = s["vB"][None, sl] * sR
    return np.ascontiguousarray(h), np.ascontiguousarray(u), np.ascontiguousarray(v)


def get_initial_conditions(ic_type: ICType, latlon_grid: LatLonGrid):
    """Return ``h, u, v, f`` on the full ghosted (M+3, N) grid (reference ic.py:22-159).

    ``h, u, v`` are freshly allocated C-contiguous float64 arrays.  For the
    Rossby-Haurwitz wave ``f`` depends on latitude only and is returned as a 2-D view of
    that vector (a real array on small grids).
    """
    g = EARTH_CONSTANTS.g
    shape = latlon_grid.shape
    f_lat = coriolis_latitude(latlon_grid)

    if ic_type == ICType.RossbyHaurwitzWave:
        s = separable_rossby_haurwitz(latlon_grid)
        cR, c2R, sR = s["cosR"][:, None], s["cos2R"][:, None], s["sinR"][:, None]
        h = s["h0"] + (s["hA"][None, :] + s["hB"][None, :] * cR + s["hC"][None, :] * c2R) / g
        u = s["uA"][None, :] + s["uB"][None, :] * cR
        v = s["vB"][None, :] * sR
        f = np.broadcast_to(f_lat[None, :], shape)
        if latlon_grid._real:
            f = np.array(f)
    elif ic_type == ICType.ZonalGeostrophicFlow:
        a, om = EARTH_CONSTANTS.a, EARTH_CONSTANTS.omega
        alpha = math.pi / 2
        u0 = 2.0 * math.pi * a / (12.0 * 24.0 * 3600.0)
        h0 = 2.94e4 / g
        th, ph = latlon_grid.theta1d[None, :], latlon_grid.phi1d[:, None]
        # ic.py:126-133: the rotated Coriolis parameter is genuinely 2-D
        rot = -np.cos(ph) * np.cos(th) * np.sin(alpha) + np.sin(th) * np.cos(alpha)
        f = 2.0 * om * rot
        h = h0 - (a * om * u0 + 0.5 * (u0 ** 2.0)) * (rot ** 2.0) / g
        u = u0 * (np.cos(th) * np.cos(alpha) + np.cos(ph) * np.sin(th) * np.sin(alpha))
        v = np.array(np.broadcast_to(-u0 * np.sin(ph) * np.sin(alpha), shape))
    else:
        raise ValueError(f"Unknown ICType: {ic_type}")

    return np.ascontiguousarray(h), np.ascontiguousarray(u), np.ascontiguousarray(v), f
