"""Dev script: adversarial inputs through the host build of the kernel logic (cold path and tracked path)
against the oracle (NumPy port of the reference) and the 40-digit truth."""
import os, sys
import numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tools')
import check_track as ct, check_emu as ce
from oracle import cassan_oracle as o
from oracle.truth_mp import truth_point
ce.lib = ct.libs

def check(name, s, q, zeta, rho=1e-3, G=0.5, tracked_seg=0):
    zeta = np.asarray(zeta, dtype=complex)
    if tracked_seg:
        ni, A, fl, _ = ct.strip(s, q, rho, G, zeta, tracked_seg)
        A0, A2, A4 = A[:, 0], A[:, 1], A[:, 2]
    else:
        ni, A0, A2, A4, fl = ce.emu_points(s, q, rho, G, zeta)
    nbad = 0; worst = 0.0; nref = 0
    for i, z in enumerate(zeta):
        try:
            t = truth_point(s, q, rho, G, z)
        except Exception as e:           # mpmath.polyroots gives up on degenerate polynomials (zeta = 0, -s)
            print(f"  {name}: zeta={z} n={ni[i]} A0={A0[i]} flags={fl[i]} reference {o.point_a4(s, q, rho, G, z)[:2]} (no truth: {type(e).__name__})")
            continue
        if ni[i] != t[0]:
            nbad += 1
            print(f"  {name}: zeta={z} n={ni[i]} truth n={t[0]} flags={fl[i]}")
            continue
        with np.errstate(all='ignore'):
            rel = max(abs(A0[i] / t[1] - 1), abs(A2[i] / t[2] - 1), abs(A4[i] / t[3] - 1)) if t[1] else 0.0
        r = o.point_a4(s, q, rho, G, z)
        nref += (r[0] != t[0])
        if not (fl[i] & (4 | 16)):
            worst = max(worst, rel)
    print(f"{name:34s} s={s:g} q={q:g} n={len(zeta)} count!=truth {nbad}  (reference itself != truth: {nref})  max rel vs truth (not near-critical) {worst:.2e}  flags {sorted(set(fl.tolist()))}")

rng = np.random.default_rng(0)
x = np.linspace(-2.5, 2.5, 60)
import warnings; warnings.simplefilter("ignore")
check("source on the lens axis", 1.0, 0.5, x + 0j)
check("source on the lens axis, tracked", 1.0, 0.5, np.linspace(-2.5, 2.5, 300) + 0j, tracked_seg=64)
check("axis, small q", 1.2, 1e-3, x + 0j)
check("just off the axis", 0.8, 0.1, x + 1e-12j)
check("far source", 1.0, 0.5, np.array([1e2, 1e3 + 1e3j, -1e4j, 1e6 + 1j]))
check("equal masses", 1.0, 1.0, rng.uniform(-1, 1, 40) + 1j * rng.uniform(-1, 1, 40))
check("close binary s=0.1", 0.1, 0.3, rng.uniform(-0.5, 0.5, 40) + 1j * rng.uniform(-0.5, 0.5, 40))
check("wide binary s=10", 10.0, 0.3, np.concatenate([rng.uniform(-1, 1, 20) + 1j * rng.uniform(-1, 1, 20), -10 + rng.uniform(-1, 1, 20) + 1j * rng.uniform(-1, 1, 20)]))
check("q=1e-5 near the planet", 1.3, 1e-5, -1.3 + 0.53 + rng.uniform(-0.02, 0.02, 40) + 1j * rng.uniform(-0.02, 0.02, 40))
check("q=1e-7", 1.0, 1e-7, rng.uniform(-0.3, 0.3, 30) + 1j * rng.uniform(-0.3, 0.3, 30))
check("near the lens positions", 1.0, 0.5, np.array([1e-6, 1e-9j, -1 + 1e-7, -1 - 1e-10j, 1e-300, -1 + 0j, 0j]))
# crossing the caustic along a line with fine steps, tracked
check("caustic crossing, tracked", 1.0, 0.5, np.linspace(-0.7, -0.2, 400) + 0.05j - 1 / 3, tracked_seg=64)
check("cusp approach, tracked", 0.8, 0.1, np.linspace(0.2, 0.6, 400) + 1e-4j - 0.8 * 0.1 / 1.1, tracked_seg=64)
