"""Dev script: warm-started segments (host emulation of k_models / k_map control flow) vs cold solves on
random C5-like models (log-uniform s, q, rho) -- image counts must match, values agree to 1e-9."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import sys
import numpy as np
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tools')
import check_track as ct, check_emu as ce
from microlensing_b200.batched import lightcurve_coordinates
ce.lib = ct.libs


def run(M=200, T=2000, L=32, seed=1, qlo=1e-4, quadratic=1, slo=0.5, shi=2.0, u0max=0.5):
    ct.libs.emu_set_quadratic(quadratic)      # k_models keeps three levels of root history
    rng = np.random.default_rng(seed)
    s = np.exp(rng.uniform(np.log(slo), np.log(shi), M))
    q = np.exp(rng.uniform(np.log(qlo), np.log(1.), M))
    rho = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    u0, al = rng.uniform(-u0max, u0max, M), rng.uniform(0, 2 * np.pi, M)
    nbad = 0; worst = 0.0; cold = 0; tot = 0; nflag = 0
    for m in range(M):
        zeta = lightcurve_coordinates(s[m], q[m], u0[m], al[m], -2.0, 4.0 / T, T)
        ni, A, fl, nc = ct.strip(s[m], q[m], rho[m], 0.5, zeta, L)
        n0, B0, B2, B4, f0 = ce.emu_points(s[m], q[m], rho[m], 0.5, zeta)
        B = np.stack([B0, B2, B4], 1)
        bad = np.where(ni != n0)[0]
        if len(bad):
            nbad += len(bad)
            print('model', m, 's,q,rho,u0,al =', s[m], q[m], rho[m], u0[m], al[m], 'nimg mismatch at', bad[:10], ni[bad][:10], n0[bad][:10])
        ok = (ni == n0) & (fl == 0) & (f0 == 0)
        with np.errstate(all='ignore'):
            rel = np.abs(A / B - 1).max(axis=1)
        if ok.any():
            w = rel[ok].max()
            if w > 1e-9:
                i = np.where(ok & (rel > 1e-9))[0]
                print('model', m, 'rel', w, 'at', i[:5], 'q', q[m])
            worst = max(worst, w)
        cold += nc; tot += T; nflag += ((fl & 1) != 0).sum()
    print(f'M={M} T={T} L={L}: nimg mismatches {nbad}, max rel (unflagged) {worst:.2e}, cold fraction {cold / tot:.4f}, noconv flags {nflag}')
    return nbad, worst


if __name__ == "__main__":
    run(M=int(sys.argv[1]) if len(sys.argv) > 1 else 200, T=int(sys.argv[2]) if len(sys.argv) > 2 else 2000,
        L=int(sys.argv[3]) if len(sys.argv) > 3 else 32)
