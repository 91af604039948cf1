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


if __name__ == "__main__":
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
