"""GPU parity tests: every call goes through the C ABI of libmlmultipoles.so (ctypes) and is compared
with the oracle / the golden vectors of the live reference.  Tolerance: 1e-9 relative on A0/A2/A4
(BASELINE.json north_star), image count exact, rules of tests/parity.py."""
import io
import contextlib

import os

import numpy as np
import pytest

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb(built):
    import microlensing_b200 as m
    m._capi.context()           # fails loudly without a GPU / without the CUDA library
    return m


def match_root_sets(a, b):
    """max distance after greedy nearest-neighbour pairing of two unordered root sets"""
    b = list(b)
    worst = 0.0
    for z in a:
        d = [abs(z - w) for w in b]
        k = int(np.argmin(d))
        worst = max(worst, d[k])
        b.pop(k)
    return worst


def truth_fn_factory(par, zeta):
    from oracle.truth_mp import truth_point
    return lambda i: truth_point(*par[i], zeta[i])


def test_example_prints_reference_values(mb):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        assert mb.multipoles.example() is None
    got = buf.getvalue().splitlines()
    ref = open(os.path.join(ROOT, "tests", "golden", "example_stdout.txt")).read().splitlines()
    assert len(got) == 4 and got[1] == ref[1] and got[3] == ref[3]
    assert got[0].startswith(" A0, A2 =  ") and got[2].startswith(" A0, A2, A4 =  ")
    for a, b in ((got[0], ref[0]), (got[2], ref[2])):
        va = [float(x) for x in a.split("=")[1].split()]
        vb = [float(x) for x in b.split("=")[1].split()]
        assert va == pytest.approx(vb, rel=1e-9)
    # the module is reachable the way the reference README says
    from microlensing import multipoles
    assert multipoles.example is mb.multipoles.example


def test_golden_points(mb, golden_points):
    g = golden_points
    par, zeta = g["params"], g["zeta"]
    n = len(zeta)
    got_n, got_A, flags = np.zeros(n, int), np.zeros((n, 3)), np.zeros(n, int)
    got_q = np.zeros((n, 2))
    keys = {}
    for i, p in enumerate(map(tuple, par)):
        keys.setdefault(p, []).append(i)
    for p, idx in keys.items():
        idx = np.array(idx)
        r = mb.magnification_points(*p, zeta[idx])
        got_n[idx], flags[idx] = r["nimg"], r["flags"]
        got_A[idx] = np.stack([r["A0"], r["A2"], r["A4"]], 1)
        r2 = mb.magnification_points(*p, zeta[idx], order=2)
        got_q[idx] = np.stack([r2["A0"], r2["A2"]], 1)
        assert np.array_equal(r2["nimg"], r["nimg"]) and not r2["A4"].any()
    mask = parity.mask_from_reference(g["hexadecapole"], g["residuals"], g["minJ"])
    rep = parity.compare(got_n, got_A, g["nimg"], g["hexadecapole"], mask, truth_fn=truth_fn_factory(par, zeta),
                         got_flags=flags)
    print("golden parity:", rep)
    assert (flags & 1).sum() == 0
    # quadrupole path == first two outputs of the hexadecapole path (bit-identical in the reference)
    assert np.array_equal(got_q, got_A[:, :2])
    # against truth: rounding level, image counts exact
    assert np.array_equal(got_n, g["truth_nimg"])
    with np.errstate(all="ignore"):
        rel = np.abs(got_A / g["truth"] - 1)
    assert np.nanmax(rel) < 1e-11


def test_flag_semantics_against_oracle(mb):
    """SURVEY.md §8(f).2 / VERDICT r1 item 4: NEARCRIT <=> oracle minJ < 1e-3, SERIES <=> |A4-A2| > |A2-A0| on oracle
    values, AMBIG / EXTRA <=> oracle-side residuals of the solver's own roots (tests/flag_semantics.py)."""
    import flag_semantics
    from oracle import cassan_oracle as oracle

    def points(s, q, rho, G, zeta):
        # every second entry is far away, so that each real position follows a jump and is solved cold
        z2 = np.empty(2 * len(zeta), dtype=np.complex128)
        z2[0::2], z2[1::2] = zeta, 50 + 50j
        r = mb.magnification_points(s, q, rho, G, z2)
        return r["nimg"][0::2], np.stack([r["A0"], r["A2"], r["A4"]], 1)[0::2], r["flags"][0::2]

    def roots(s, q, zeta):
        ctx = mb._capi.context()
        zeta = np.ascontiguousarray(zeta, dtype=np.complex128)
        out = np.empty((len(zeta), 5), dtype=np.complex128)
        ctx.check(ctx.lib.mlm_solve_host(ctx.handle, s, q, mb._capi.ptr(zeta), len(zeta), mb._capi.ptr(out)), "mlm_solve_host")
        return out

    flag_semantics.check(points, roots, oracle)


def test_direct_call_drop_ins(mb, golden_calls):
    c = golden_calls
    mp = mb.multipoles
    off = 0
    for k, n in enumerate(c["wk_n"]):
        Wk = c["wk_flat"][off:off + 7 * n].reshape(7, n)
        off += 7 * n
        q2 = mp.quadrupole(Wk, c["wk_rho"][k], c["wk_gamma"][k])
        h4 = mp.hexadecapole(Wk, c["wk_rho"][k], c["wk_gamma"][k])
        assert q2.shape == (2,) and h4.shape == (3,) and q2.dtype == np.float64
        assert np.allclose(q2, c["quad_out"][k], rtol=1e-9, atol=0), (k, n)
        assert np.allclose(h4, c["hex_out"][k], rtol=1e-9, atol=0), (k, n)
        if n == 0:
            assert not h4.any() and not q2.any()
    # rows 0-1 are never read: garbage there must not matter (np.empty in the reference, multipoles.py:186)
    Wk = c["wk_flat"][:7 * 0 + 0]                              # (first entry has n=0); build a fresh one
    rng = np.random.default_rng(3)
    Wk = np.zeros((7, 5), complex)
    Wk[2:] = 0.3 * (rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5)))
    a = mp.hexadecapole(Wk, 0.01, 0.3)
    Wk[:2] = np.nan
    assert np.array_equal(mp.hexadecapole(Wk, 0.01, 0.3), a)
    # a (7, 2, 3) table sums over every trailing axis (axis-less np.sum, multipoles.py:144-146)
    assert np.allclose(mp.hexadecapole(Wk[:, :4].reshape(7, 2, 2), 0.01, 0.3),
                       mp.hexadecapole(np.ascontiguousarray(Wk[:, :4]), 0.01, 0.3), rtol=1e-15)
    with pytest.raises(IndexError):
        mp.hexadecapole(np.zeros((5, 3), complex), 0.01, 0.)
    # _akl
    got = mp._akl(c["akl_mu"], c["akl_W2"], c["akl_Q"])
    assert got.shape == c["akl_out"].shape
    assert np.allclose(got, c["akl_out"], rtol=1e-14, atol=0)
    assert np.allclose(mp._akl(2.0, 0.5 + 0.1j, 1 - 2j), 2.0 * (np.conj(1 - 2j) + np.conj(0.5 + 0.1j) * (1 - 2j)))
    # _solvelenseq: same root set as np.roots (order unspecified)
    for (s, q, z), ref in zip(c["solve_in"], c["solve_out"]):
        got = mp._solvelenseq(s.real, q.real, z)
        assert got.shape == (5,) and got.dtype == np.complex128
        assert match_root_sets(got, ref) < 1e-9
    got = mp._solvelenseq(1.7, 0.2, np.complex128(-0.7833333333333334))
    assert len(got) == 5
    with pytest.raises(ValueError):
        mp._solvelenseq(1.0, 0.5, np.array([0.1 + 0.1j, 0.2]))
    assert len(mp._solvelenseq(1.0, 0.5, 0j)) == 4 and len(mp._solvelenseq(1.0, 0.5, -1 + 0j)) == 4


@pytest.mark.parametrize("name,s,q,cx,half,nx", [("C4", 1.0, 0.5, -0.08, 0.80, 384), ("C3", 0.8, 0.1, 0.11, 0.75, 320)])
def test_map_vs_batch_oracle(mb, name, s, q, cx, half, nx):
    from oracle import cassan_oracle as oracle
    rho, G = 1e-3, 0.5
    ny = nx
    x0, y0, dx, dy = cx - half, -half, 2 * half / nx, 2 * half / ny
    m = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny)
    zeta = mb.map_coordinates(s, q, x0, y0, dx, dy, nx, ny)
    nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
    ref_A = np.stack([B0.ravel(), B2.ravel(), B4.ravel()], 1)
    mask = parity.mask_from_reference(ref_A, res, minJ.ravel())
    got_A = np.stack([m["A0"].ravel(), m["A2"].ravel(), m["A4"].ravel()], 1)
    zf = zeta.ravel()
    par = [(s, q, rho, G)] * zf.size
    rep = parity.compare(m["nimg"].ravel(), got_A, nb.ravel(), ref_A, mask, truth_fn=truth_fn_factory(par, zf),
                         got_flags=m["flags"].ravel())
    print(name, "map parity:", rep, "f5 =", float((nb == 5).mean()))
    assert rep["masked_frac"] < 0.01
    assert (m["flags"] & 1).sum() == 0
    # points API (cold root solve per position) vs the map kernel (warm-started tracking down the
    # columns): same roots up to the Newton stopping rule -> identical image counts, values at
    # rounding level (amplified next to caustics, where the mask applies)
    p = mb.magnification_points(s, q, rho, G, zeta)
    assert np.array_equal(p["nimg"], m["nimg"])
    for k in ("A0", "A2", "A4"):
        rel = np.abs(p[k] / m[k] - 1)
        assert rel[~mask.reshape(rel.shape)].max() < 1e-10, k


def test_lightcurve_C2_vs_oracle(mb):
    from oracle import cassan_oracle as oracle
    s, q, rho, G = 1.0, 1e-3, 1e-3, 0.5
    T = 20000
    zeta = mb.lightcurve_coordinates(s, q, 0.01, 0.35, -2.0, 4.0 / T, T)
    r = mb.magnification_points(s, q, rho, G, zeta)
    nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
    ref_A = np.stack([B0, B2, B4], 1)
    mask = parity.mask_from_reference(ref_A, res, minJ)
    got_A = np.stack([r["A0"], r["A2"], r["A4"]], 1)
    par = [(s, q, rho, G)] * T
    rep = parity.compare(r["nimg"], got_A, nb, ref_A, mask, truth_fn=truth_fn_factory(par, zeta), got_flags=r["flags"])
    print("C2 parity:", rep)
    # models API with M=1 (in-kernel trajectory, warm-started tracking along the epochs) gives the
    # same light curve: same parity rules against the reference, rounding-level agreement with the
    # cold-solved points away from the masked positions
    mm = mb.magnification_models([s], [q], [rho], [G], [0.01], [0.35], -2.0, 4.0 / T, T)
    got_M = np.stack([mm["A0"][0], mm["A2"][0], mm["A4"][0]], 1)
    rep2 = parity.compare(mm["nimg"][0], got_M, nb, ref_A, mask, truth_fn=truth_fn_factory(par, zeta),
                          got_flags=mm["flags"][0])
    print("C2 parity (tracked light curve):", rep2)
    assert np.array_equal(mm["nimg"][0], r["nimg"])
    rel = np.abs(got_M / got_A - 1).max(axis=1)
    assert rel[~mask].max() < 1e-9 and np.median(rel) < 1e-13
    assert (mm["flags"] & 3).sum() == 0


@pytest.mark.parametrize("s,q", [(0.8, 0.1), (1.0, 1e-3)])
def test_points_on_an_uneven_curved_trajectory(mb, s, q):
    """magnification_points walks consecutive list entries with warm-started tracking (extrapolation
    scaled by the ratio of successive source steps): unevenly sampled, curved trajectory with gaps and a
    shuffled copy of the same positions (no coherence at all) against the oracle."""
    from oracle import cassan_oracle as oracle
    rho, G = 1e-3, 0.5
    rng = np.random.default_rng(3)
    n = 30000
    tau = np.cumsum(rng.exponential(1.0, n))
    tau = -1.5 + 3.0 * tau / tau[-1]                                   # uneven epochs
    tau[n // 3:] += 0.2                                                # a gap (a jump -> cold restart)
    zeta_cm = (tau + 1j * (0.02 + 0.05 * tau ** 2)) * np.exp(0.3j)      # curved
    zeta = zeta_cm - s * q / (1 + q)
    nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
    ref_A = np.stack([B0, B2, B4], 1)
    mask = parity.mask_from_reference(ref_A, res, minJ)
    par = [(s, q, rho, G)] * n
    r = mb.magnification_points(s, q, rho, G, zeta)
    rep = parity.compare(r["nimg"], np.stack([r["A0"], r["A2"], r["A4"]], 1), nb, ref_A, mask,
                         truth_fn=truth_fn_factory(par, zeta), got_flags=r["flags"])
    print("uneven trajectory parity:", rep)
    assert (r["flags"] & 1).sum() == 0
    perm = rng.permutation(n)
    rs = mb.magnification_points(s, q, rho, G, zeta[perm])
    assert np.array_equal(rs["nimg"], r["nimg"][perm])
    ok = (rs["flags"] == 0) & (r["flags"][perm] == 0)
    assert np.abs(rs["A4"] / r["A4"][perm] - 1)[ok].max() < 1e-9


def test_unordered_list_entries_are_solved_independently(mb):
    """mlm_points_unordered (what magnification_points picks for a list without spatial order): every entry solved on
    its own -- three single-lens image guesses + Newton on the lens equation, the root finder on the compacted rest --
    results at the entries' own places, independent of the order of the list.  Against the tracking kernel on the same
    list (one cold solve per entry there), and against the oracle on a sample."""
    import torch
    from oracle import cassan_oracle as oracle
    from microlensing_b200.batched import points_device
    rng = np.random.default_rng(2718)
    # a window on a resonant caustic (about half of the entries go to the second pass), a planetary configuration
    # (nearly all solved by the guesses) and a wide field
    for (s, q, rho, G, n, win) in ((1.0, 0.5, 1e-3, 0.5, 60000, 0.3), (0.9, 1e-3, 1e-3, 0.5, 40000, 0.15),
                                   (1.0, 0.5, 1e-3, 0.5, 20000, 2.0)):
        zeta = rng.uniform(-win, win, n) + 1j * rng.uniform(-win, win, n)
        zeta[17] = zeta[4711]                                          # duplicates are fine
        auto = mb.magnification_points(s, q, rho, G, zeta)            # >= 8192 entries without order -> mlm_points_unordered
        dev = torch.device("cuda", torch.cuda.current_device())
        zt = torch.from_numpy(zeta).to(dev)
        b = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)] + \
            [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
        st = torch.cuda.current_stream().cuda_stream
        res = {}
        for unordered in (False, True):
            for t in b:
                t.fill_(0)
            points_device(s, q, rho, G, zt, n, *b, stream=st, unordered=unordered)
            torch.cuda.synchronize()
            res[unordered] = [t.cpu().numpy().copy() for t in b]
        plain, srt = res[False], res[True]
        assert np.array_equal(auto["nimg"], srt[3]) and np.array_equal(auto["A4"], srt[2])      # the host API took that path
        assert np.array_equal(plain[3], srt[3])                        # image counts: tracked == cold
        ok = (plain[4] == 0) & (srt[4] == 0)
        for k in range(3):
            assert np.abs(srt[k][ok] / plain[k][ok] - 1).max() < 1e-9, k
        assert (srt[4] & 1).sum() == 0
        idx = rng.choice(n, 4000, replace=False)
        nb, B0, B2, B4, z, keep, minJ, resid = oracle.batch_a4(s, q, rho, G, zeta[idx], return_extra=True)
        ref_A = np.stack([B0, B2, B4], 1)
        rep = parity.compare(srt[3][idx], np.stack([srt[0][idx], srt[1][idx], srt[2][idx]], 1), nb, ref_A,
                             parity.mask_from_reference(ref_A, resid, minJ), got_flags=srt[4][idx],
                             truth_fn=truth_fn_factory([(s, q, rho, G)] * len(idx), zeta[idx]))
        print("unordered list parity:", rep)
        # order-independence: the same entries in another order give the same bits at their new places
        perm = rng.permutation(n)
        zt2 = torch.from_numpy(zeta[perm]).to(dev)
        points_device(s, q, rho, G, zt2, n, *b, stream=st, unordered=True)
        torch.cuda.synchronize()
        for k in range(5):
            assert np.array_equal(b[k].cpu().numpy(), srt[k][perm]), k
    # non-finite entries do not disturb the others
    zeta = rng.uniform(-1, 1, 9000) + 1j * rng.uniform(-1, 1, 9000)
    good = mb.magnification_points(1.0, 0.5, 1e-3, 0.5, zeta)
    zeta2 = zeta.copy()
    zeta2[5] = complex(np.nan, 0.0)
    bad = mb.magnification_points(1.0, 0.5, 1e-3, 0.5, zeta2)
    keep = np.arange(9000) != 5
    assert np.array_equal(bad["nimg"][keep], good["nimg"][keep])
    ok = keep & (good["flags"] == 0) & (bad["flags"] == 0)
    assert np.abs(bad["A4"][ok] / good["A4"][ok] - 1).max() < 1e-9


def test_models_C5_sample_vs_oracle(mb):
    from oracle import cassan_oracle as oracle
    rng = np.random.default_rng(20240517)
    M, T = 24, 400
    s = np.exp(rng.uniform(np.log(0.5), np.log(2.), M))
    q = np.exp(rng.uniform(np.log(1e-4), np.log(1.), M))
    rho = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    G = np.full(M, 0.5)
    u0, al = rng.uniform(-.5, .5, M), rng.uniform(0, 2 * np.pi, M)
    r = mb.magnification_models(s, q, rho, G, u0, al, -2.0, 4.0 / T, T)
    assert r["A4"].shape == (M, T)
    got_n, got_A, ref_n, ref_A, masks, par, zs = [], [], [], [], [], [], []
    for m in range(M):
        zeta = mb.lightcurve_coordinates(s[m], q[m], u0[m], al[m], -2.0, 4.0 / T, T)
        nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s[m], q[m], rho[m], G[m], zeta, return_extra=True)
        ra = np.stack([B0, B2, B4], 1)
        ref_n.append(nb); ref_A.append(ra); masks.append(parity.mask_from_reference(ra, res, minJ))
        got_n.append(r["nimg"][m]); got_A.append(np.stack([r["A0"][m], r["A2"][m], r["A4"][m]], 1))
        par += [(s[m], q[m], rho[m], G[m])] * T
        zs.append(zeta)
    zs = np.concatenate(zs)
    rep = parity.compare(np.concatenate(got_n), np.concatenate(got_A), np.concatenate(ref_n), np.concatenate(ref_A),
                         np.concatenate(masks), truth_fn=truth_fn_factory(par, zs), got_flags=r["flags"].ravel())
    print("C5 parity:", rep)
    assert rep["nimg_mismatch"] <= 0.005 * rep["n"]      # 11 of 9600 on the host-emulated logic, all at q = 1.6e-4


def test_models_chi2_fused_fit(mb):
    """the fused chi^2 kernel equals the unfused path (light curves from
    magnification_models, sums in NumPy) and, for one model, the oracle's light curve."""
    from oracle import cassan_oracle as oracle
    rng = np.random.default_rng(99)
    M, T = 20, 700
    s = np.exp(rng.uniform(np.log(0.6), np.log(1.6), M))
    q = np.exp(rng.uniform(np.log(1e-3), np.log(1.), M))
    rho = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    G = np.full(M, 0.5)
    u0, al = rng.uniform(-.4, .4, M), rng.uniform(0, 2 * np.pi, M)
    tau0, dtau = -2.0, 4.0 / T
    lc = mb.magnification_models(s, q, rho, G, u0, al, tau0, dtau, T)
    sigma = 0.01 + 0.002 * rng.random(T)
    flux = 3.0 * lc["A4"][0] + 0.7 + sigma * rng.normal(size=T)        # model 0 is the truth: Fs = 3, Fb = 0.7
    r = mb.models_chi2(s, q, rho, G, u0, al, tau0, dtau, flux, sigma)
    w = 1.0 / sigma ** 2

    def fit(A):
        S, SA, SAA, SF, SAF, SFF = w.sum(), (w * A).sum(), (w * A * A).sum(), (w * flux).sum(), (w * A * flux).sum(), (w * flux * flux).sum()
        det = S * SAA - SA * SA
        Fs, Fb = (S * SAF - SA * SF) / det, (SAA * SF - SA * SAF) / det
        return ((w * (flux - Fs * A - Fb) ** 2).sum(), Fs, Fb, SA, SAA, SAF)

    for m in range(M):
        chi2, Fs, Fb, SA, SAA, SAF = fit(lc["A4"][m])
        assert np.allclose([r["Fs"][m], r["Fb"][m], r["SA"][m], r["SAA"][m], r["SAF"][m]], [Fs, Fb, SA, SAA, SAF],
                           rtol=1e-10), m
        assert abs(r["chi2"][m] - chi2) <= 1e-7 * max(chi2, 1.0), (m, r["chi2"][m], chi2)   # chi2 = SFF - Fs SAF - Fb SF cancels
        assert r["n5"][m] == (lc["nimg"][m] == 5).sum() and r["nflag"][m] == (lc["flags"][m] != 0).sum()
    assert abs(r["Fs"][0] - 3.0) < 0.01 and abs(r["Fb"][0] - 0.7) < 0.02
    assert r["chi2"][0] == r["chi2"].min() and 0.8 * T < r["chi2"][0] < 1.2 * T
    # one model against the oracle's light curve (reference formulae)
    m = 1
    zeta = mb.lightcurve_coordinates(s[m], q[m], u0[m], al[m], tau0, dtau, T)
    nb, B0, B2, B4 = oracle.batch_a4(s[m], q[m], rho[m], G[m], zeta)[:4]
    chi2, Fs, Fb = fit(B4)[:3]
    assert np.allclose([r["Fs"][m], r["Fb"][m]], [Fs, Fb], rtol=1e-7)
    assert abs(r["chi2"][m] - chi2) <= 1e-6 * chi2
    # order 2 uses the quadrupole light curve
    r2 = mb.models_chi2(s, q, rho, G, u0, al, tau0, dtau, flux, sigma, order=2)
    assert np.allclose(r2["SA"], (w[None, :] * lc["A2"]).sum(axis=1), rtol=1e-10)
    # long segments: 700 epochs in segments of 100 -> 8 threads per model, 8 models share a block of 64 threads
    # (what a 1e4-model grid gets); same light curves to rounding, same statistics from them
    assert mb._capi.context().auto_segment(10000, 2000) == 32 and mb._capi.context().auto_segment(1, 1000000) <= 16
    for seg in (100, 350, 700):
        lcs = mb.magnification_models(s, q, rho, G, u0, al, tau0, dtau, T, segment=seg)
        assert np.array_equal(lcs["nimg"], lc["nimg"]), seg
        ok = (lcs["flags"] == 0) & (lc["flags"] == 0)
        assert np.abs(lcs["A4"] / lc["A4"] - 1)[ok].max() < 1e-9, seg
        rs = mb.models_chi2(s, q, rho, G, u0, al, tau0, dtau, flux, sigma, segment=seg)
        for m in range(M):
            chi2, Fs, Fb, SA, SAA, SAF = fit(lcs["A4"][m])
            assert np.allclose([rs["Fs"][m], rs["Fb"][m], rs["SA"][m], rs["SAA"][m], rs["SAF"][m]], [Fs, Fb, SA, SAA, SAF],
                               rtol=1e-10), (seg, m)
            assert rs["n5"][m] == (lcs["nimg"][m] == 5).sum() and rs["nflag"][m] == (lcs["flags"][m] != 0).sum()


def test_models_at_observation_times(mb):
    """Light curves at unevenly spaced observation times with per-model (t0, tE) -- the in-kernel trajectory
    zeta(t; t0, tE, u0, alpha) of SURVEY.md §8(f).1 -- against the point solver on the same positions and,
    for the fused fit, against NumPy sums over those light curves."""
    rng = np.random.default_rng(5)
    M, T = 12, 900
    s = np.exp(rng.uniform(np.log(0.6), np.log(1.6), M))
    q = np.exp(rng.uniform(np.log(1e-3), np.log(1.), M))
    rho = np.full(M, 1e-3)
    G = np.full(M, 0.5)
    u0, al = rng.uniform(-.3, .3, M), rng.uniform(0, 2 * np.pi, M)
    t0 = 2459000.0 + rng.uniform(-5, 5, M)
    tE = rng.uniform(15, 60, M)
    # nightly clusters of exposures with gaps (days), like a ground-based survey
    nights = 2459000.0 - 40 + np.sort(rng.choice(80, 60, replace=False)).astype(float)
    times = np.sort(np.concatenate([n + 0.3 + np.sort(rng.uniform(0, 0.3, T // 60)) for n in nights]))
    T = times.size
    r = mb.magnification_models(s, q, rho, G, u0, al, times=times, t0=t0, tE=tE)
    assert r["A4"].shape == (M, T)
    for m in range(M):
        zeta = mb.lightcurve_coordinates_at(s[m], q[m], u0[m], al[m], times, t0[m], tE[m])
        p = mb.magnification_points(s[m], q[m], rho[m], G[m], zeta)
        assert np.array_equal(p["nimg"], r["nimg"][m]), m
        ok = (p["flags"] == 0) & (r["flags"][m] == 0)
        for k in ("A0", "A2", "A4"):
            assert np.abs(r[k][m] / p[k] - 1)[ok].max() < 1e-9, (m, k)
    assert ((r["flags"] & 1) == 0).all()
    # scalar t0 / tE and the default (times already in Einstein times) agree with the uniform-grid kernel
    tau = -2.0 + 4.0 / 500 * np.arange(500)
    a = mb.magnification_models(s, q, rho, G, u0, al, times=tau)
    b = mb.magnification_models(s, q, rho, G, u0, al, -2.0, 4.0 / 500, 500)
    assert np.array_equal(a["nimg"], b["nimg"])
    okab = (a["flags"] == 0) & (b["flags"] == 0)
    assert np.abs(a["A4"] / b["A4"] - 1)[okab].max() < 1e-9
    # fused fit at observation times
    sigma = np.full(T, 0.02)
    flux = 1.5 * r["A4"][3] + 0.2 + sigma * rng.normal(size=T)
    c = mb.models_chi2(s, q, rho, G, u0, al, flux=flux, sigma=sigma, times=times, t0=t0, tE=tE)
    w = 1 / sigma ** 2
    for m in range(M):
        A = r["A4"][m]
        assert np.allclose([c["SA"][m], c["SAA"][m], c["SAF"][m]], [(w * A).sum(), (w * A * A).sum(), (w * A * flux).sum()],
                           rtol=1e-10)
    assert int(np.argmin(c["chi2"])) == 3 and abs(c["Fs"][3] - 1.5) < 0.05


def test_image_count_is_odd_where_the_reference_loses_images(mb):
    """Images almost on top of a lens (very distant source, q <= 1e-5): the reference counts 2, the truth and
    the kernel 3; values equal the 40-digit evaluation."""
    from oracle.truth_mp import truth_point
    rng = np.random.default_rng(11)
    cases = [(1.0, 0.5, np.array([1e2 + 0j, 1e3 + 1e3j, -1e4j, 1e6 + 1j])),
             (1.0, 1e-7, rng.uniform(-0.3, 0.3, 12) + 1j * rng.uniform(-0.3, 0.3, 12)),
             (1.3, 1e-5, -1.3 + 0.53 + rng.uniform(-0.02, 0.02, 12) + 1j * rng.uniform(-0.02, 0.02, 12))]
    for s, q, zeta in cases:
        r = mb.magnification_points(s, q, 1e-3, 0.5, zeta)
        for i, z in enumerate(zeta):
            t = truth_point(s, q, 1e-3, 0.5, z)
            assert r["nimg"][i] == t[0], (s, q, z, r["nimg"][i], t[0])
            assert np.allclose([r["A0"][i], r["A2"][i], r["A4"][i]], t[1:4], rtol=1e-12), (s, q, z)


def test_edge_cases(mb, golden_points):
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    z = np.array([0j, -1 + 0j, 1e-9 + 0j, 1e-6 + 1e-6j, -1 + 1e-8j, 50 + 20j])
    r = mb.magnification_points(s, q, rho, G, z)
    g = golden_points
    m = np.where(g["tags"] == "edge")[0][:6]
    assert np.array_equal(r["nimg"], g["nimg"][m])            # 3 images each, like the reference
    assert np.allclose(np.stack([r["A0"], r["A2"], r["A4"]], 1), g["hexadecapole"][m], rtol=1e-9)
    assert (r["flags"][:2] & 32).all()                         # degenerate polynomial reported
    # empty and odd-sized inputs
    e = mb.magnification_points(s, q, rho, G, np.zeros(0, complex))
    assert e["A0"].shape == (0,)
    one = mb.magnification_points(s, q, rho, G, 0.1 + 0.2j)
    assert one["A4"].shape == () and one["nimg"][()] in (3, 5)
    odd = mb.magnification_points(s, q, rho, G, np.full((3, 43), 0.1 + 0.2j))
    assert odd["A4"].shape == (3, 43) and np.all(odd["A4"] == one["A4"][()])
    em = mb.magnification_map(s, q, rho, G, -1., -1., 0.1, 0.1, 0, 0)
    assert em["A0"].shape == (0, 0)
    # argument errors surface as ValueError (negative return codes of the C ABI)
    with pytest.raises(ValueError):
        mb.magnification_points(-1.0, q, rho, G, z)
    with pytest.raises(ValueError):
        mb.magnification_points(s, q, rho, G, z, order=3)
    # NaN input does not hang or crash: zero images
    nn = mb.magnification_points(s, q, rho, G, np.array([complex(np.nan, 0.0)]))
    assert nn["nimg"][0] == 0 and nn["A0"][0] == 0.0


def test_row_sharding_is_bitwise_identical(mb):
    """sharded results must equal single-GPU results bitwise."""
    from microlensing_b200.sharding import shard_range
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    nx, ny = 256, 203
    x0, y0, dx, dy = -0.88, -0.8, 1.6 / nx, 1.6 / ny
    from microlensing_b200.sharding import auto_strip
    full = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny)
    strip = auto_strip(nx, ny)                      # what the host API picks for this map, whatever the row block
    assert strip == 16 and mb._capi.context().auto_strip(nx, ny) == 16   # same rule in the C ABI (mlm_auto_strip)
    assert mb._capi.context().auto_strip(8192, 8192) == 64 and mb._capi.context().auto_strip(8192, 1024) == 16
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_range(ny, r, world, align=strip)        # shards start on strip boundaries
            parts.append(mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=lo, nrows=hi - lo))
        for k in ("A0", "A2", "A4", "nimg", "flags"):
            assert np.array_equal(np.concatenate([p[k] for p in parts], 0), full[k]), (world, k)
    # cyclic sharding: strips k = r (mod world) written by "rank" r into the same full-map buffers
    import torch
    from microlensing_b200.batched import map_strips_device
    for world in (2, 3):
        bufs = [torch.full((ny * nx,), -1.0, dtype=torch.float64, device="cuda") for _ in range(3)] + \
               [torch.full((ny * nx,), 255, dtype=torch.uint8, device="cuda") for _ in range(2)]
        for r in range(world):                           # the strip length is an argument of the call
            map_strips_device(s, q, rho, G, x0, y0, dx, dy, nx, 0, ny, r, world, *bufs, strip=strip)
        torch.cuda.synchronize()
        for k, b in zip(("A0", "A2", "A4", "nimg", "flags"), bufs):
            assert np.array_equal(b.cpu().numpy().reshape(ny, nx), full[k]), (world, k)
    # weighted deals (the gather root owns more strips), also on a row window that starts and ends inside strips
    from microlensing_b200.batched import map_deal_device
    from microlensing_b200.sharding import deal_rounds
    for world, share, row0, nrows in ((2, 0.6, 0, ny), (4, 0.31, 0, ny), (8, 0.3, 0, ny), (3, 0.5, 21, 150)):
        period, sel = deal_rounds(world, share)
        bufs = [torch.full((nrows * nx,), -1.0, dtype=torch.float64, device="cuda") for _ in range(3)] + \
               [torch.full((nrows * nx,), 255, dtype=torch.uint8, device="cuda") for _ in range(2)]
        for r in range(world):
            map_deal_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, nrows, period, sel[r], *bufs, strip=strip)
        torch.cuda.synchronize()
        ref = full if nrows == ny else mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=row0, nrows=nrows)
        for k, b in zip(("A0", "A2", "A4", "nimg", "flags"), bufs):
            assert np.array_equal(b.cpu().numpy().reshape(nrows, nx), ref[k]), (world, share, k)
    # unaligned shards still agree to rounding (a strip cut in two starts with one more cold solve)
    a = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=0, nrows=50)
    b = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=50, nrows=ny - 50)
    assert np.array_equal(np.concatenate([a["nimg"], b["nimg"]], 0), full["nimg"])
    assert np.allclose(np.concatenate([a["A4"], b["A4"]], 0), full["A4"], rtol=1e-9, atol=0)


def test_device_pointer_api_and_sharded_map(mb):
    import torch
    from microlensing_b200.sharding import ShardedMap
    s, q, rho, G = 0.8, 0.1, 1e-3, 0.5
    nx = ny = 192
    x0, y0, dx, dy = 0.11 - 0.75, -0.75, 1.5 / nx, 1.5 / ny
    from microlensing_b200.sharding import auto_strip
    strip = auto_strip(nx, ny)                       # = what magnification_map picks for this map
    sm = ShardedMap(nx, ny, 0, 1, nchunks=3, align=strip)
    sm.compute(s, q, rho, G, x0, y0, dx, dy)
    torch.cuda.synchronize()
    dev = sm.assemble_host()
    host = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny)
    for k in ("A0", "A2", "A4", "nimg", "flags"):
        assert np.array_equal(dev[k], host[k]), k
    # output subset: unselected arrays are never written
    sm2 = ShardedMap(nx, ny, 0, 1, outputs=("A4", "flags"), align=strip)
    for k in ("A0", "A2", "nimg"):
        sm2.local[k].fill_(7)
    sm2.compute(s, q, rho, G, x0, y0, dx, dy)
    torch.cuda.synchronize()
    d2 = sm2.assemble_host()
    assert np.array_equal(d2["A4"], host["A4"]) and np.array_equal(d2["flags"], host["flags"])
    assert (d2["A0"] == 7).all() and (d2["nimg"] == 7).all()


def test_host_chunking_is_bitwise_neutral(mb):
    """The host entry point pipelines a large map in row chunks (kernel k+1 overlaps D2H k); the chunks
    are whole strips, so the result equals the single-launch device API bit for bit -- also for an
    odd width and a row range that starts inside a strip."""
    import torch
    from microlensing_b200.batched import map_device
    s, q, rho, G = 0.8, 0.1, 1e-3, 0.5
    nx, ny, row0 = 1000, 4300, 37
    x0, y0, dx, dy = 0.11 - 0.75, -0.75, 1.5 / nx, 1.5 / ny
    nrows = ny - row0
    host = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, row0=row0, nrows=nrows)
    n = nrows * nx
    bufs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)] + \
           [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
    from microlensing_b200.sharding import auto_strip
    map_device(s, q, rho, G, x0, y0, dx, dy, nx, row0, nrows, *bufs, strip=auto_strip(nx, ny))
    torch.cuda.synchronize()
    for k, b in zip(("A0", "A2", "A4", "nimg", "flags"), bufs):
        assert np.array_equal(b.cpu().numpy().reshape(nrows, nx), host[k]), k


def test_staged_peer_store_path_is_bitwise_neutral(mb):
    """The store path k_map takes when its outputs live in ANOTHER GPU's memory (one-byte arrays staged in shared memory
    and written a 128-byte row segment at a time) is forced onto local outputs and compared byte for byte with the
    plain stores: row windows starting inside a strip, partial groups of four rows, NULL outputs, a dealt launch, order 2
    (a 1-GPU box never takes that path otherwise; untouched arrays keep their fill value in both runs)."""
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_staged_stores.py")], capture_output=True,
                       text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0 and "all identical" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_output_selection(mb):
    """NULL outputs in the C ABI: unselected arrays are neither stored nor copied; the others are unchanged."""
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    nx = ny = 96
    x0, y0, dx, dy = -0.88, -0.8, 1.6 / nx, 1.6 / ny
    full = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny)
    part = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, outputs=("A4", "nimg"))
    assert set(part) == {"A4", "nimg"}
    assert np.array_equal(part["A4"], full["A4"]) and np.array_equal(part["nimg"], full["nimg"])
    zeta = mb.map_coordinates(s, q, x0, y0, dx, dy, nx, ny)
    pp = mb.magnification_points(s, q, rho, G, zeta, outputs=("A0", "flags"))
    pf = mb.magnification_points(s, q, rho, G, zeta)
    assert np.array_equal(pp["A0"], pf["A0"]) and np.array_equal(pp["flags"], pf["flags"])
    with pytest.raises(ValueError):
        mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, nx, ny, outputs=())


def _two_gpu_worker(rank, world, port, gather, layout, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MLM_DEVICE"] = str(rank)
    if layout == "fallback":                       # p2p requested, peer mapping "unavailable" -> NCCL gather
        os.environ["MLM_DISABLE_IPC"] = "1"
        layout = None
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from microlensing_b200.sharding import ShardedMap
        s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
        nx, ny = 512, 320
        x0, y0, dx, dy = -0.88, -0.8, 1.6 / nx, 1.6 / ny
        sm = ShardedMap(nx, ny, rank, world, gather=gather, layout=layout, align=8,
                        root_share=0.6 if layout == "weighted" else None)
        if os.environ.get("MLM_DISABLE_IPC"):
            assert sm.gather == "nccl" and sm.layout == "block"
        if layout == "weighted":
            assert sm.period == 5 and len(sm.sel_all[0]) == 3
        sm.compute(s, q, rho, G, x0, y0, dx, dy, local_only=True)     # kernel-only leg: must not disturb the result
        torch.cuda.synchronize()
        if sm.stage is not None:                  # poison the staging map: what reaches rank 0 must come from THIS run
            for v in sm.stage.values():
                v.fill_(float("nan") if v.dtype == torch.float64 else 255)
        if rank == 0 and getattr(sm, "full_map", None) is not None:
            for v in sm.full_map.values():
                v.fill_(float("nan") if v.dtype == torch.float64 else 255)
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(2):
            sm.compute(s, q, rho, G, x0, y0, dx, dy)
        torch.cuda.synchronize()
        full = sm.assemble_host()
        if rank == 0:
            np.savez(os.path.join(out_dir, "full.npz"), **full)
        dist.barrier()
        sm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather,layout", [("p2p", "cyclic"), ("p2p", "weighted"), ("bulk", "weighted"), ("bulk", "cyclic"),
                                           ("p2p", "block"), ("nccl", "block"), ("p2p", "fallback")])
def test_two_gpu_sharded_map_equals_single_gpu(mb, tmp_path, gather, layout):
    """Row-sharded over 2 GPUs with the fused peer-store gather (and the NCCL gather): rank 0's
    assembled map is bit-identical to the single-GPU map."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_two_gpu_worker, args=(2, port, gather, layout, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "full.npz")
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    nx, ny = 512, 320
    # the single-GPU map with the strip length the sharded run picked (auto_strip) is bit-identical;
    # with the default strip it agrees to rounding
    from microlensing_b200.sharding import auto_strip
    default = mb.magnification_map(s, q, rho, G, -0.88, -0.8, 1.6 / nx, 1.6 / ny, nx, ny)
    assert auto_strip(nx, ny // 2) == 16 and auto_strip(8192, 8192) == 64 and auto_strip(8192, 1024) == 16
    one = mb.magnification_map(s, q, rho, G, -0.88, -0.8, 1.6 / nx, 1.6 / ny, nx, ny, strip=8)   # the workers' align=8
    for k in ("A0", "A2", "A4", "nimg", "flags"):
        assert np.array_equal(got[k], one[k]), k
    assert np.array_equal(got["nimg"], default["nimg"])
    ok = (default["flags"] == 0) & (got["flags"] == 0)
    assert np.abs(got["A4"] / default["A4"] - 1)[ok].max() < 1e-9


def _two_gpu_models_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MLM_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from microlensing_b200.sharding import ShardedModels
        par, T, flux, sigma = _model_grid_case()
        sm = ShardedModels(par, -2.0, 4.0 / T, T, rank, world)
        sm.compute()
        res = sm.result_host()
        sf = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, fit=(flux, sigma))
        sf.compute()
        fit = sf.result_host()
        s64 = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, segment=64)      # 8 threads per model, 8 models per block
        s64.compute()
        res64 = s64.result_host()
        f64 = ShardedModels(par, -2.0, 4.0 / T, T, rank, world, fit=(flux, sigma), segment=64)
        f64.compute()
        fit64 = f64.result_host()
        s64.close(); f64.close()
        par8, times = _model_grid_times()
        st = ShardedModels(par8, rank=rank, world=world, times=times)
        st.compute()
        rt = st.result_host()
        if rank == 0:
            np.savez(os.path.join(out_dir, "models.npz"), **res, **{"fit_" + k: v for k, v in fit.items()},
                     times_A4=rt["A4"], times_nimg=rt["nimg"], **{"s64_" + k: v for k, v in res64.items()},
                     **{"f64_" + k: v for k, v in fit64.items()})
        dist.barrier()
        sm.close()
        sf.close()
        st.close()
    finally:
        dist.destroy_process_group()


def _model_grid_case():
    rng = np.random.default_rng(77)
    M, T = 37, 500
    s = np.exp(rng.uniform(np.log(0.5), np.log(2.), M))
    q = np.exp(rng.uniform(np.log(1e-4), np.log(1.), M))
    rho = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    par = (s, q, rho, np.full(M, 0.5), rng.uniform(-.5, .5, M), rng.uniform(0, 2 * np.pi, M))
    flux = 1.0 + rng.random(T)
    return par, T, flux, np.full(T, 0.05)


def _model_grid_times():
    par, T, _, _ = _model_grid_case()
    rng = np.random.default_rng(78)
    M = par[0].size
    times = np.sort(rng.uniform(-30.0, 30.0, 333))
    return par + (rng.uniform(-2, 2, M), rng.uniform(15, 40, M)), times


def test_sharded_models_single_and_two_gpu(mb, tmp_path):
    """Model-sharded grid (BASELINE configs[4]): ShardedModels on 1 rank, and on 2 ranks when 2 GPUs are
    visible, equals the plain host API bit for bit (magnifications and fused fit statistics)."""
    import torch
    from microlensing_b200.sharding import ShardedModels
    par, T, flux, sigma = _model_grid_case()
    one = mb.magnification_models(*par, -2.0, 4.0 / T, T)
    fit1 = mb.models_chi2(*par, -2.0, 4.0 / T, flux, sigma)
    sm = ShardedModels(par, -2.0, 4.0 / T, T)
    sm.compute()
    res = sm.result_host()
    sf = ShardedModels(par, -2.0, 4.0 / T, T, fit=(flux, sigma))
    sf.compute()
    fit = sf.result_host()
    sm.close(); sf.close()
    for k in ("A0", "A2", "A4", "nimg", "flags"):
        assert np.array_equal(res[k], one[k]), k
    for k in ShardedModels.STAT_KEYS:
        assert np.array_equal(fit[k], fit1[k]), k
    # a light curve does not depend on how many models share its call (host chunks, shards): model 5 alone, same segment
    for seg in (4, 64):
        full = mb.magnification_models(*par, -2.0, 4.0 / T, T, segment=seg)
        alone = mb.magnification_models(*[a[5:6] for a in par], -2.0, 4.0 / T, T, segment=seg)
        for k in ("A0", "A2", "A4", "nimg", "flags"):
            assert np.array_equal(alone[k][0], full[k][5]), (seg, k)
    par8, times = _model_grid_times()
    onet = mb.magnification_models(*par8[:6], times=times, t0=par8[6], tE=par8[7])
    stt = ShardedModels(par8, times=times)
    stt.compute()
    rt = stt.result_host()
    stt.close()
    assert np.array_equal(rt["A4"], onet["A4"]) and np.array_equal(rt["nimg"], onet["nimg"])
    if torch.cuda.device_count() < 2:
        return
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_two_gpu_models_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "models.npz")
    for k in ("A0", "A2", "A4", "nimg", "flags"):
        assert np.array_equal(got[k], one[k]), k
    for k in ShardedModels.STAT_KEYS:
        assert np.array_equal(got["fit_" + k], fit1[k]), k
    assert np.array_equal(got["times_A4"], onet["A4"]) and np.array_equal(got["times_nimg"], onet["nimg"])
    # ... and with long segments (several models per block), for which the sharded ranks see different model counts
    one64 = mb.magnification_models(*par, -2.0, 4.0 / T, T, segment=64)
    fit64 = mb.models_chi2(*par, -2.0, 4.0 / T, flux, sigma, segment=64)
    for k in ("A0", "A2", "A4", "nimg", "flags"):
        assert np.array_equal(got["s64_" + k], one64[k]), k
    for k in ShardedModels.STAT_KEYS:
        assert np.array_equal(got["f64_" + k], fit64[k]), k


def test_full_size_properties(mb):
    """BASELINE sizes through size-independent properties: the lens is symmetric about the real
    axis, so a map whose window is symmetric in y equals its own vertical mirror; image counts are
    3 or 5; A0 >= 1 (total point-source magnification of a binary lens)."""
    s, q, rho, G = 1.0, 0.5, 1e-3, 0.5
    nx, ny = 8192, 1024                      # 8192-wide rows of the headline map, both central bands
    half = 0.8125                            # exactly representable: rows r and 8191-r are exact mirrors in y
    x0, dx, dy = -0.08 - half, 2 * half / 8192, 2 * half / 8192
    top = mb.magnification_map(s, q, rho, G, x0, -half, dx, dy, nx, 8192, row0=4096 - ny, nrows=ny)
    bot = mb.magnification_map(s, q, rho, G, x0, -half, dx, dy, nx, 8192, row0=4096, nrows=ny)
    assert np.array_equal(top["nimg"], bot["nimg"][::-1])
    # the two bands are walked in opposite directions by the warm-started tracking, so the roots
    # differ by rounding; next to caustics that rounding is amplified (flagged pixels)
    clean = (top["flags"] == 0) & (bot["flags"][::-1] == 0)
    assert clean.mean() > 0.99
    for k in ("A0", "A2", "A4"):
        rel = np.abs(top[k] / bot[k][::-1] - 1)
        assert rel[clean].max() < 1e-9, (k, rel[clean].max())
        assert np.median(rel) < 1e-14, k
    assert set(np.unique(top["nimg"])) <= {3, 5}
    assert top["A0"].min() >= 1.0
    assert ((top["flags"] & 3) == 0).all()
    f5 = float((top["nimg"] == 5).mean())
    assert 0.05 < f5 < 0.6


# ---- BASELINE.json configurations at FULL size -----------------------------------------------------
def _sample_map_vs_oracle(mb, m, s, q, rho, G, x0, y0, dx, dy, nx, ny, nsample, seed):
    from oracle import cassan_oracle as oracle
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, nx * ny, nsample)
    r, c = np.divmod(idx, nx)
    zeta = ((x0 + (c + 0.5) * dx) + (-(s * q / (1. + q)))) + 1j * (y0 + (r + 0.5) * dy)
    nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
    ref_A = np.stack([B0, B2, B4], 1)
    mask = parity.mask_from_reference(ref_A, res, minJ)
    got_A = np.stack([m["A0"].ravel()[idx], m["A2"].ravel()[idx], m["A4"].ravel()[idx]], 1)
    par = [(s, q, rho, G)] * nsample
    return parity.compare(m["nimg"].ravel()[idx], got_A, nb, ref_A, mask, truth_fn=truth_fn_factory(par, zeta),
                          got_flags=m["flags"].ravel()[idx])


@pytest.mark.parametrize("name,s,q,cx,half,n", [("C4", 1.0, 0.5, -0.08, 0.80, 8192), ("C3", 0.8, 0.1, 0.11, 0.75, 4096)])
def test_full_size_maps_sampled_against_oracle(mb, name, s, q, cx, half, n):
    """The whole 8192^2 / 4096^2 map of BASELINE.json through the public host API; 40 000 random pixels
    of it against the oracle (same parity rules), plus whole-map invariants."""
    rho, G = 1e-3, 0.5
    x0, y0, dx, dy = cx - half, -half, 2 * half / n, 2 * half / n
    m = mb.magnification_map(s, q, rho, G, x0, y0, dx, dy, n, n)
    rep = _sample_map_vs_oracle(mb, m, s, q, rho, G, x0, y0, dx, dy, n, n, 40000, 2024)
    print(name, "full-size sampled parity:", rep)
    assert rep["masked_frac"] < 0.01
    assert set(np.unique(m["nimg"])) <= {3, 5}
    assert ((m["flags"] & 3) == 0).all()                    # no iteration cap hit anywhere on the map
    assert m["A0"].min() >= 1.0 and np.isfinite(m["A4"]).all()
    assert (m["flags"] != 0).mean() < 0.01


def test_full_size_C2_lightcurve(mb):
    """1e6 epochs (BASELINE configs[1]): tracked light-curve kernel vs cold-solved points on every epoch,
    and 30 000 of them against the oracle."""
    from oracle import cassan_oracle as oracle
    s, q, rho, G = 1.0, 1e-3, 1e-3, 0.5
    T = 1000000
    dtau = 4.0 / T
    zeta = mb.lightcurve_coordinates(s, q, 0.01, 0.35, -2.0, dtau, T)
    p = mb.magnification_points(s, q, rho, G, zeta)
    mm = mb.magnification_models([s], [q], [rho], [G], [0.01], [0.35], -2.0, dtau, T)
    assert np.array_equal(mm["nimg"][0], p["nimg"])
    ok = (mm["flags"][0] == 0) & (p["flags"] == 0)
    assert ok.mean() > 0.99
    for k in ("A0", "A2", "A4"):
        assert np.abs(mm[k][0] / p[k] - 1)[ok].max() < 1e-9, k
    assert ((mm["flags"] & 3) == 0).all() and ((p["flags"] & 3) == 0).all()
    idx = np.sort(np.random.default_rng(5).integers(0, T, 30000))
    nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s, q, rho, G, zeta[idx], return_extra=True)
    ref_A = np.stack([B0, B2, B4], 1)
    got_A = np.stack([mm["A0"][0][idx], mm["A2"][0][idx], mm["A4"][0][idx]], 1)
    rep = parity.compare(mm["nimg"][0][idx], got_A, nb, ref_A, parity.mask_from_reference(ref_A, res, minJ),
                         truth_fn=truth_fn_factory([(s, q, rho, G)] * len(idx), zeta[idx]), got_flags=mm["flags"][0][idx])
    print("C2 full-size parity:", rep)


def test_full_size_C5_model_grid(mb):
    """1e4 models x 2000 epochs (BASELINE configs[4]): whole-grid invariants, 40 models against the cold
    point solver and 6 models against the oracle."""
    from oracle import cassan_oracle as oracle
    rng = np.random.default_rng(20240517)
    M, T = 10000, 2000
    s = np.exp(rng.uniform(np.log(0.5), np.log(2.), M))
    q = np.exp(rng.uniform(np.log(1e-4), np.log(1.), M))
    rho = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), M))
    G = np.full(M, 0.5)
    u0, al = rng.uniform(-.5, .5, M), rng.uniform(0, 2 * np.pi, M)
    r = mb.magnification_models(s, q, rho, G, u0, al, -2.0, 4.0 / T, T)
    assert set(np.unique(r["nimg"])) <= {3, 5}
    assert ((r["flags"] & 1) == 0).all()
    assert r["A0"].min() >= 1.0 and np.isfinite(r["A4"]).all()
    pick = rng.choice(M, 40, replace=False)
    agg = []
    for j, m in enumerate(pick):
        zeta = mb.lightcurve_coordinates(s[m], q[m], u0[m], al[m], -2.0, 4.0 / T, T)
        p = mb.magnification_points(s[m], q[m], rho[m], G[m], zeta)
        assert np.array_equal(p["nimg"], r["nimg"][m]), m
        ok = (p["flags"] == 0) & (r["flags"][m] == 0)
        assert np.abs(r["A4"][m] / p["A4"] - 1)[ok].max() < 1e-9, m
        if j < 6:
            nb, B0, B2, B4, z, keep, minJ, res = oracle.batch_a4(s[m], q[m], rho[m], G[m], zeta, return_extra=True)
            ref_A = np.stack([B0, B2, B4], 1)
            agg.append((r["nimg"][m], np.stack([r["A0"][m], r["A2"][m], r["A4"][m]], 1), nb, ref_A,
                        parity.mask_from_reference(ref_A, res, minJ), r["flags"][m], zeta,
                        [(s[m], q[m], rho[m], G[m])] * T))
    cat = [np.concatenate([a[i] for a in agg]) for i in range(7)]
    rep = parity.compare(cat[0], cat[1], cat[2], cat[3], cat[4], got_flags=cat[5],
                         truth_fn=truth_fn_factory(sum((a[7] for a in agg), []), cat[6]))
    print("C5 full-size parity (6 models x 2000 epochs):", rep)
