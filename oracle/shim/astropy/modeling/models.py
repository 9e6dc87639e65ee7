import numpy as np

_h = 6.62607015e-27
_c = 2.99792458e10
_kB = 1.380649e-16


class BlackBody:
    """B_nu(T) in erg s^-1 cm^-2 Hz^-1 sr^-1, as astropy.modeling.models.BlackBody(T)(nu)."""

    def __init__(self, temperature):
        self.T = float(temperature)

    def __call__(self, nu):
        nu = np.asarray(nu, dtype=np.float64)
        with np.errstate(over="ignore"):
            return 2.0 * _h * nu ** 3 / _c ** 2 / np.expm1(_h * nu / (_kB * self.T))
