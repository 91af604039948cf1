"""Small run of every kernel through the host API (no torch), for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import microlensing_b200 as mb
from microlensing_b200.multipoles import multipoles as mp

mb.multipoles.example()
s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
m = mb.magnification_map(s, q, rho, G, -0.88, -0.8, 1.6 / 200, 1.6 / 131, 200, 131)          # odd sizes, partial blocks
p = mb.magnification_points(s, q, rho, G, mb.map_coordinates(s, q, -0.88, -0.8, 1.6 / 200, 1.6 / 131, 200, 131))
assert np.array_equal(m["nimg"], p["nimg"])
part = mb.magnification_map(s, q, rho, G, -0.88, -0.8, 1.6 / 200, 1.6 / 131, 200, 131, row0=70, nrows=33, outputs=("A4",))
assert np.array_equal(part["A4"], m["A4"][70:103]) or np.allclose(part["A4"], m["A4"][70:103], rtol=1e-9)
r = mb.magnification_models([1.0, 0.7, 1.3], [1e-3, 0.1, 1e-4], [1e-3] * 3, [0.5] * 3, [0.01, 0.2, -0.1], [0.35, 1.0, 2.0], -2.0, 4.0 / 333, 333)
c = mb.models_chi2([1.0, 0.7, 1.3], [1e-3, 0.1, 1e-4], [1e-3] * 3, [0.5] * 3, [0.01, 0.2, -0.1], [0.35, 1.0, 2.0], -2.0, 4.0 / 333,
                   2.0 * r["A4"][0] + 0.3, np.full(333, 0.01))
assert int(np.argmin(c["chi2"])) == 0
z = np.array([0.3 + 0.2j, -0.5 + 0.1j, 1.2 - 0.7j])
Wk = np.zeros((7, 3), complex)
for k in range(2, 7):
    Wk[k] = (-1) ** (k - 1) * np.prod(np.arange(1, k)) / (1 + q) * (z ** -k + q * (z + s) ** -k)
print(mp.quadrupole(Wk, rho, G), mp.hexadecapole(Wk, rho, G), mp._solvelenseq(s, q, 0.1 + 0.2j), mp._akl(1.0, Wk[2], Wk[3]))
print("sanitize_small OK; launches:", mb._capi.context().launch_count())
