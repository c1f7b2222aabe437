"""Constants and type aliases of the shallow-water path.

Mirrors the names of the reference's ``src/sw_solver/utils.py:7-36`` (``FloatT``,
``EARTH_CONSTANTS``); the numerical values are part of the parity contract.
"""

from dataclasses import dataclass

import numpy as np

FloatT = np.float64

FloatArray1D = np.ndarray
FloatArray2D = np.ndarray
FloatArray3D = np.ndarray


@dataclass(frozen=True)
class EarthConstants:
    """Physical constants of the planet (reference utils.py:15-33).

    g [m/s2] gravity, a [m] radius, omega [1/s] rotation rate, nu [m2/s] viscosity;
    rho and scaleHeight are carried for API compatibility only (unused by the path).
    """

    g: float = 9.80616
    rho: float = 1.2
    a: float = 6.37122e6
    omega: float = 7.292e-5
    scaleHeight: float = 8.0e3
    nu: float = 5.0e5


EARTH_CONSTANTS = EarthConstants()
