import numpy as np


class _Unit:
    """A unit that only knows its CGS scale; `x * unit` returns x unchanged (plain float64)."""
    __array_ufunc__ = None      # make ndarray * unit defer to __rmul__

    def __init__(self, name, scale):
        self.name = name
        self.scale = scale

    def to(self, other):
        return self.scale / other.scale

    def __rmul__(self, x):
        return np.asarray(x, dtype=np.float64) if isinstance(x, np.ndarray) else x

    __mul__ = __rmul__


s = _Unit("s", 1.0)
yr = _Unit("yr", 3.15576e7)
cm = _Unit("cm", 1.0)
pc = _Unit("pc", 3.0856775814913674e18)
kpc = _Unit("kpc", 3.0856775814913674e21)
g = _Unit("g", 1.0)
M_p = _Unit("M_p", 1.67262192369e-24)
M_e = _Unit("M_e", 9.1093837015e-28)
erg = _Unit("erg", 1.0)
eV = _Unit("eV", 1.602176634e-12)
K = _Unit("K", 1.0)
Hz = _Unit("Hz", 1.0)
