#!/usr/bin/env python
"""Generate golden vectors by running the LIVE, UNMODIFIED reference.

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``/root/reference/microlensing/multipoles/multipoles.py`` (with the
caller-side shim ``np.complex = complex`` that NumPy >= 1.24 needs,
SURVEY.md §0.4), replays ``multipoles.py:181-202`` per position and stores
inputs, reference outputs and the 40-digit "truth" of the same formulae
(``oracle/truth_mp.py``) in ``tests/golden/golden_points.npz``, plus direct
``quadrupole/hexadecapole/_akl/_solvelenseq`` call fixtures in
``tests/golden/golden_calls.npz``.  NumPy version and seeds are recorded.
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
np.complex = complex  # shim for multipoles.py:177-178 on NumPy >= 1.24

from microlensing.multipoles import multipoles as ref  # noqa: E402  (the live reference)
from oracle.truth_mp import truth_point  # noqa: E402


def ref_point(s, q, rho, Gamma, zeta):
    """Literal replay of multipoles.py:181-202 using the reference's own functions."""
    zeta = np.complex128(zeta)
    with np.errstate(all="ignore"):
        z_all = ref._solvelenseq(s, q, zeta)
        W1 = 1. / (1. + q) * (1. / z_all + q / (z_all + s))
        res = np.abs(z_all - W1.conjugate() - zeta)
        z0 = z_all[res < 0.000001]
        nr = len(z0)
        Wk = np.zeros((7, nr), dtype=np.complex128)
        Wk[2] = -1. / (1. + q) * (1. / z0**2 + q / (z0 + s)**2)
        Wk[3] = 2. / (1. + q) * (1. / z0**3 + q / (z0 + s)**3)
        Wk[4] = -6. / (1. + q) * (1. / z0**4 + q / (z0 + s)**4)
        Q = ref.quadrupole(Wk, rho, Gamma)
        Wk[5] = 24. / (1. + q) * (1. / z0**5 + q / (z0 + s)**5)
        Wk[6] = -120. / (1. + q) * (1. / z0**6 + q / (z0 + s)**6)
        H = ref.hexadecapole(Wk, rho, Gamma)
        minJ = np.min(np.abs(1. - np.abs(Wk[2])**2)) if nr else np.inf
    roots = np.full(5, np.nan + 1j * np.nan)
    roots[:len(z_all)] = z_all
    resid = np.full(5, np.nan)
    resid[:len(z_all)] = res
    return nr, Q, H, roots, resid, minJ


def build_cases():
    """(tag, s, q, rho, Gamma, zeta_primary) tuples; windows from SURVEY.md §8(d)."""
    cases = []
    # Appendix B table (primary-frame zeta)
    appB = [
        ("appB", 1.7, 0.2, 0.01, 0.0, -0.7833333333333334 + 0j),
        ("appB", 1.7, 0.2, 0.01, 0.5, -0.7833333333333334 + 0j),
        ("appB", 1.0, 0.5, 0.001, 0.5, -0.2333333333333333 + 0.05j),
        ("appB", 1.0, 0.5, 0.001, 0.5, 0.26666666666666666 - 0.45j),
        ("appB", 0.8, 0.1, 0.001, 0.5, -0.052727272727272734 + 0.01j),
        ("appB", 0.8, 0.1, 0.001, 0.5, -0.5727272727272728 + 0.3j),
        ("appB", 1.0, 0.001, 0.001, 0.5, -0.030999000999001 + 0.004j),
        ("appB", 1.0, 0.001, 0.001, 0.5, 0.8990009990009991 + 0.02j),
    ]
    cases += appB
    rng = np.random.default_rng(20261019)
    # C2: light curve s=1, q=1e-3, u0=0.01, alpha=0.35, tau in [-2,2]
    s, q, rho, G = 1.0, 1e-3, 1e-3, 0.5
    tau = np.sort(rng.uniform(-2, 2, 700))
    tau = np.concatenate([tau, rng.uniform(-0.2, 0.2, 500)])       # oversample the caustic region
    zcm = (tau + 1j * 0.01) * np.exp(1j * 0.35)
    cases += [("C2", s, q, rho, G, z - s * q / (1 + q)) for z in zcm]
    # C3: 4096^2 map window s=0.8 q=0.1 centre (0.11,0) half-width 0.75
    s, q = 0.8, 0.1
    for _ in range(1500):
        z = complex(0.11 + rng.uniform(-.75, .75), rng.uniform(-.75, .75))
        cases.append(("C3", s, q, rho, G, z - s * q / (1 + q)))
    # C4: 8192^2 map window s=1 q=0.5 centre (-0.08,0) half-width 0.80
    s, q = 1.0, 0.5
    for _ in range(2000):
        z = complex(-0.08 + rng.uniform(-.8, .8), rng.uniform(-.8, .8))
        cases.append(("C4", s, q, rho, G, z - s * q / (1 + q)))
    # C5: random models x a few epochs each
    for _ in range(400):
        s = float(np.exp(rng.uniform(np.log(0.5), np.log(2.))))
        q = float(np.exp(rng.uniform(np.log(1e-4), np.log(1.))))
        rho = float(np.exp(rng.uniform(np.log(1e-4), np.log(1e-2))))
        u0, al = rng.uniform(-.5, .5), rng.uniform(0, 2 * np.pi)
        for tau in rng.uniform(-2, 2, 3):
            z = (tau + 1j * u0) * np.exp(1j * al)
            cases.append(("C5", s, q, rho, 0.5, z - s * q / (1 + q)))
    # edge cases of SURVEY App. A.4 (s=1, q=0.5)
    for z in (0j, -1.0 + 0j, 1e-9 + 0j, 1e-6 + 1e-6j, -1.0 + 1e-8j, 50. + 20j, -300j):
        cases.append(("edge", 1.0, 0.5, 1e-3, 0.5, z))
    return cases


def main():
    cases = build_cases()
    n = len(cases)
    tags = np.array([c[0] for c in cases])
    par = np.array([[c[1], c[2], c[3], c[4]] for c in cases], dtype=np.float64)
    zeta = np.array([c[5] for c in cases], dtype=np.complex128)
    nimg = np.zeros(n, np.int64)
    quad = np.zeros((n, 2))
    hexa = np.zeros((n, 3))
    roots = np.zeros((n, 5), np.complex128)
    resid = np.zeros((n, 5))
    minJ = np.zeros(n)
    t_n = np.zeros(n, np.int64)
    truth = np.zeros((n, 3))
    for i, (tag, s, q, rho, G, z) in enumerate(cases):
        nimg[i], quad[i], hexa[i], roots[i], resid[i], minJ[i] = ref_point(s, q, rho, G, z)
        t = truth_point(s, q, rho, G, z)
        t_n[i], truth[i] = t[0], t[1:]
        if i % 500 == 0:
            print(i, n, flush=True)
    np.savez_compressed(
        os.path.join(HERE, "golden_points.npz"),
        tags=tags, params=par, zeta=zeta, nimg=nimg, quadrupole=quad, hexadecapole=hexa,
        roots=roots, residuals=resid, minJ=minJ, truth_nimg=t_n, truth=truth,
        numpy_version=np.array(np.__version__), seed=np.array(20261019))

    # example() printed output, verbatim
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.example()
    with open(os.path.join(HERE, "example_stdout.txt"), "w") as f:
        f.write(buf.getvalue())

    # direct-call fixtures: quadrupole/hexadecapole on arbitrary Wk, _akl, _solvelenseq
    rng = np.random.default_rng(7)
    wk_list, quad_out, hex_out = [], [], []
    for nr in (0, 1, 2, 3, 5, 8, 64):
        for _ in range(6):
            Wk = np.zeros((7, nr), np.complex128)
            # keep |W2| away from 1, mixed magnitudes
            Wk[2:] = (rng.normal(size=(5, nr)) + 1j * rng.normal(size=(5, nr))) * \
                np.array([0.4, 1., 2., 5., 20.])[:, None]
            rho, G = float(rng.uniform(1e-4, 5e-2)), float(rng.uniform(0, 1))
            wk_list.append((Wk, rho, G))
            with np.errstate(all="ignore"):
                quad_out.append(ref.quadrupole(Wk, rho, G))
                hex_out.append(ref.hexadecapole(Wk, rho, G))
    mu = rng.normal(size=100)
    W2 = rng.normal(size=100) + 1j * rng.normal(size=100)
    Qk = rng.normal(size=100) + 1j * rng.normal(size=100)
    akl = ref._akl(mu, W2, Qk)
    solve_in = [(1.7, 0.2, -0.7833333333333334 + 0j), (1.0, 0.5, 0.1 + 0.2j), (0.8, 0.1, -0.3 - 0.4j),
                (1.0, 1e-3, 0.01 + 0.002j), (2.0, 1e-4, 1.5 + 0.1j), (0.5, 1.0, -0.2 + 0.0j)]
    solve_out = [ref._solvelenseq(s, q, np.complex128(z)) for s, q, z in solve_in]
    np.savez_compressed(
        os.path.join(HERE, "golden_calls.npz"),
        wk_n=np.array([w[0].shape[1] for w in wk_list]),
        wk_flat=np.concatenate([w[0].ravel() for w in wk_list]),
        wk_rho=np.array([w[1] for w in wk_list]), wk_gamma=np.array([w[2] for w in wk_list]),
        quad_out=np.array(quad_out), hex_out=np.array(hex_out),
        akl_mu=mu, akl_W2=W2, akl_Q=Qk, akl_out=akl,
        solve_in=np.array(solve_in, dtype=np.complex128), solve_out=np.array(solve_out))
    print("wrote golden fixtures:", n, "points")


if __name__ == "__main__":
    main()
