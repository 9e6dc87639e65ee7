"""CGS constants exactly as the reference obtains them from astropy 5.0.4 (LeMoC/LeHaMoC_f.py:17-31)."""
import numpy as np

G = 6.6743e-8
c = 2.99792458e10
Ro = 6.957e10
Mo = 1.988409870698051e33
yr = 3.15576e7
kpc = 3.0856775814913674e21
pc = 3.0856775814913674e18
m_pr = 1.67262192369e-24
m_el = 9.1093837015e-28
kb = 1.380649e-16
h = 6.62607015e-27
q = 4.803204712570263e-10
sigmaT = 6.6524587321e-25
eV = 1.602176634e-12
B_cr = 2 * np.pi * m_el ** 2 * c ** 3 / (h * q)
SENTINEL = 10 ** (-260.)          # LeMoC.py:174-183
