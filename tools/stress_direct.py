"""Dev script: the three-image-guess cold start (point_magnification_direct, host emulation) against the root-finder
path (point_magnification) on random positions: image counts and flags must be equal, values agree to 1e-9 where
unflagged; prints how often the guesses solve the position and the evaluation counts."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import ctypes, sys
import numpy as np
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tools')
import check_emu as ce
libs = ctypes.CDLL(_ROOT + '/tests/host_emu/libmlm_emu_stats.so')
dp = ctypes.POINTER(ctypes.c_double); up = ctypes.POINTER(ctypes.c_ubyte)
NAMES = ['laguerre', 'newton', 'lens', 'fallbacks', 'cold', 'classify_calls', 'classify_newton', 'nonphys', 'images', 'quadratics']


def allstats(reset=1):
    st = np.zeros(10, np.int64); libs.emu_stats_all(st.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), reset); return st


def direct(s, q, rho, G, zeta):
    zeta = np.ascontiguousarray(zeta, dtype=np.complex128); n = zeta.size
    A = [np.zeros(n) for _ in range(3)]; ni = np.zeros(n, np.uint8); fl = np.zeros(n, np.uint8); used = np.zeros(n, np.uint8)
    libs.emu_points_direct(ctypes.c_double(s), ctypes.c_double(q), ctypes.c_double(rho), ctypes.c_double(G),
                           zeta.view(np.float64).ctypes.data_as(dp), ctypes.c_long(n), ctypes.c_int(4),
                           *[a.ctypes.data_as(dp) for a in A], ni.ctypes.data_as(up), fl.ctypes.data_as(up), used.ctypes.data_as(up))
    return ni, np.stack(A, 1), fl, used


def run(name, s, q, zeta, rho=1e-3):
    ce.lib = libs
    allstats()
    ni, A, fl, used = direct(s, q, rho, 0.5, zeta)
    st = allstats()
    n0, B0, B2, B4, f0 = ce.emu_points(s, q, rho, 0.5, zeta)
    B = np.stack([B0, B2, B4], 1)
    bad = ni != n0
    fdiff = (fl != f0) & ~bad
    ok = ~bad & (fl == 0) & (f0 == 0)
    with np.errstate(all='ignore'):
        rel = np.abs(A / B - 1).max(axis=1)
    w = rel[ok].max() if ok.any() else 0.0
    extra_d = ((fl >> 6) & 3).astype(int); extra_0 = ((f0 >> 6) & 3).astype(int)
    print(f"{name}: N={zeta.size} solved by guesses {(used != 0).mean():.4f} (with classification {(used == 2).mean():.4f})  nimg mismatch {bad.sum()}  flag differences {fdiff.sum()} "
          f"(EXTRA count differs {np.sum(extra_d != extra_0)})  max rel unflagged {w:.2e}  lens evals/pt {st[2] / zeta.size:.2f} lag/pt {st[0] / zeta.size:.2f}")
    if bad.any():
        i = np.where(bad)[0][:5]
        print('   mismatch at', zeta[i], ni[i], n0[i], used[i], fl[i], f0[i])
    return bad.sum()


def parity_accounting():
    """Rule 5 of tests/parity.py (count disputes with the reference flagged EXTRA or reference-side ambiguous) for the direct path
    and the root-finder path on the same random positions at small q: the two paths account for the disputes alike."""
    sys.path.insert(0, _ROOT + '/tests')
    import parity
    from oracle import cassan_oracle as o
    from oracle.truth_mp import truth_point
    ce.lib = libs
    rng = np.random.default_rng(3)
    for s, q in ((1.013, 8.81e-5), (1.46, 3.86e-5), (0.863, 1.91e-4), (1.0, 1e-3), (1.2, 1e-5), (0.9, 3e-4)):
        z = rng.uniform(-2, 2, 6000) + 1j * rng.uniform(-2, 2, 6000)
        ni, A, fl, used = direct(s, q, 1e-3, 0.5, z)
        n0, B0, B2, B4, f0 = ce.emu_points(s, q, 1e-3, 0.5, z)
        nb, R0, R2, R4, zz, keep, minJ, res = o.batch_a4(s, q, 1e-3, 0.5, z, return_extra=True)
        ref_A = np.stack([R0, R2, R4], 1)
        mask = parity.mask_from_reference(ref_A, res, minJ)
        for name, (gn, gA, gf) in (('direct', (ni, A, fl)), ('root finder', (n0, np.stack([B0, B2, B4], 1), f0))):
            rep = parity.compare(gn, gA, nb, ref_A, mask, truth_fn=lambda i: truth_point(s, q, 1e-3, 0.5, z[i]), got_flags=gf,
                                 strict=False, max_adjudicate=20)
            print(f"s={s} q={q:.1e} {name:11s}: AMBIG {((gf & 8) != 0).mean():.3f}  count disputes {rep['nimg_mismatch']}  flagged EXTRA "
                  f"{rep.get('nimg_mismatch_flagged_extra')}  reference-side ambiguous {rep.get('nimg_mismatch_ref_ambiguous')}  unaccounted "
                  f"{rep.get('nimg_mismatch_unaccounted')}  adjudication failures {len(rep['adjudication_failures'])}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--parity":
        parity_accounting()
        sys.exit(0)
    rng = np.random.default_rng(11)
    N = 100000
    s, q = 1.0, 1e-3; tau = rng.uniform(-2, 2, N); run('C2', s, q, (tau + 0.01j) * np.exp(0.35j) - s * q / (1 + q))
    s, q = 1.0, 0.5; run('C4', s, q, (-0.08 + rng.uniform(-.8, .8, N)) + 1j * rng.uniform(-.8, .8, N) - s * q / (1 + q))
    s, q = 0.8, 0.1; run('C3', s, q, (0.11 + rng.uniform(-.75, .75, N)) + 1j * rng.uniform(-.75, .75, N) - s * q / (1 + q))
    s, q = 1.0, 1e-3; run('C2core', s, q, rng.uniform(-.05, .05, N) + 1j * rng.uniform(-.05, .05, N) - s * q / (1 + q))
    for k in range(16):
        s = float(np.exp(rng.uniform(np.log(0.3), np.log(3)))); q = float(np.exp(rng.uniform(np.log(1e-7), 0)))
        z = rng.uniform(-2, 2, 20000) + 1j * rng.uniform(-2, 2, 20000)
        run(f'rand s={s:.3f} q={q:.2e}', s, q, z)
    # adversarial: on the axis, far away, next to the lenses
    s, q = 1.2, 1e-5
    run('axis', s, q, rng.uniform(-3, 3, 5000) + 0j)
    run('far', s, q, 1e3 * (rng.uniform(-1, 1, 5000) + 1j * rng.uniform(-1, 1, 5000)))
    run('on planet', s, q, -s + 1e-3 * (rng.uniform(-1, 1, 5000) + 1j * rng.uniform(-1, 1, 5000)))
    run('on primary', s, q, 1e-3 * (rng.uniform(-1, 1, 5000) + 1j * rng.uniform(-1, 1, 5000)))
