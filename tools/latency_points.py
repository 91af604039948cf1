import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import microlensing_b200 as mb
s1, q1, rho1, G1 = 1.7, 0.2, 0.01, 0.0
zeta0 = complex(-.5, 0.) - complex(s1 * q1 / (1. + q1), 0.)
for n in (1, 100, 4096, 5000):
    z = np.full(n, zeta0) + 1e-3*np.arange(n)
    fn = lambda: mb.magnification_points(s1, q1, rho1, G1, z)
    for _ in range(20): fn()
    t0=time.perf_counter()
    for _ in range(200): r=fn()
    print(n, 'us/call', (time.perf_counter()-t0)/200*1e6, r['A4'][0], r['nimg'][0])
