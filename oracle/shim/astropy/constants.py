class _Val:
    def __init__(self, v):
        self.value = v


class _Const:
    def __init__(self, cgs, gauss=None):
        self.cgs = _Val(cgs)
        if gauss is not None:
            self.gauss = _Val(gauss)


G = _Const(6.6743e-8)
c = _Const(2.99792458e10)
R_sun = _Const(6.957e10)
M_sun = _Const(1.988409870698051e33)
k_B = _Const(1.380649e-16)
h = _Const(6.62607015e-27)
e = _Const(4.803204712570263e-10, gauss=4.803204712570263e-10)
sigma_T = _Const(6.6524587321e-25)
