"""What the kernel's per-position flag bits MEAN, pinned to oracle-side quantities (SURVEY.md §8(f).2, §8(c) rule 3).

Shared by the CPU test (host build of the kernel logic) and the GPU test (C ABI).  `points(s, q, rho, G, zeta) ->
(nimg, A (n,3), flags)` must solve every position COLD (an unordered list), `roots(s, q, zeta) -> (n, 5)` are the
solver's polynomial roots (the _solvelenseq drop-in).

  NEARCRIT (4)  <=>  oracle minJ = min over the reference's images of |1 - |W2|^2| < 1e-3
  SERIES  (16)  <=>  |A4 - A2| > |A2 - A0| on the oracle's values
  AMBIG    (8)  <=>  a root of the solver has an oracle-side residual |z - conj(W1) - zeta| in (1e-7, 1e-5)
  EXTRA    (2), count in bits 6-7  <=>  nimg - #(solver roots with oracle-side residual < 1e-6) > 0, equal to the count
Positions within a stated band of a threshold are excluded from the equivalence (rounding decides there); the bands
and the number of positions actually exercising each flag are asserted.
"""
from __future__ import annotations

import numpy as np

NEARCRIT, SERIES, AMBIG, EXTRA, EXTRA_SHIFT = 4, 16, 8, 2, 6


def caustic_points(s, q, n, rng, dmin=1e-9, dmax=1e-4):
    """n source positions at log-uniform distance in [dmin, dmax] from the caustic (random side, primary frame)."""
    phi = rng.uniform(0, 2 * np.pi, n)
    out = np.empty(n, dtype=np.complex128)
    for i, ph in enumerate(phi):
        e = (1 + q) * np.exp(1j * ph)
        zc = np.roots([-e, -2 * s * e, (1 + q) - e * s * s, 2 * s, s * s])        # |W2| = 1
        z = zc[rng.integers(0, 4)]
        W1 = (1 / z + q / (z + s)) / (1 + q)
        d = np.exp(rng.uniform(np.log(dmin), np.log(dmax)))
        out[i] = z - np.conj(W1) + d * np.exp(1j * rng.uniform(0, 2 * np.pi))
    return out


def check(points, roots, oracle, report=print):
    rng = np.random.default_rng(424242)
    stats = dict(n=0, nearcrit=0, series=0, ambig=0, extra=0, band_excluded=0)
    cases = []
    # (s, q, rho): the BASELINE lenses near their caustics (NEARCRIT, SERIES), and small-q lenses over a wide
    # window, where the quintic's root noise reaches the 1e-6 threshold (AMBIG, EXTRA)
    for s, q, rho in ((1.0, 0.5, 1e-3), (0.8, 0.1, 1e-3), (1.0, 1e-3, 1e-3), (1.0, 0.5, 5e-2)):
        z = np.concatenate([caustic_points(s, q, 300, rng), caustic_points(s, q, 100, rng, 1e-4, 1e-1)])
        cases.append((s, q, rho, z))
    for lq in (-3.5, -4.0, -4.5, -5.0, -6.0):
        s = float(np.exp(rng.uniform(np.log(0.6), np.log(1.8))))
        z = rng.uniform(-2, 2, 400) + 1j * rng.uniform(-2, 2, 400)
        cases.append((s, 10.0 ** lq, 1e-3, z))
    for s, q, rho, zeta in cases:
        G = 0.5
        nimg, A, fl = points(s, q, rho, G, zeta)
        nimg, fl = np.asarray(nimg).astype(int), np.asarray(fl).astype(int)
        n_ref, B0, B2, B4, _, _, minJ, _ = oracle.batch_a4(s, q, rho, G, zeta, return_extra=True)
        same_images = nimg == n_ref          # minJ / series are statements about the SAME image set
        # NEARCRIT
        band = np.abs(minJ / 1e-3 - 1) < 0.02
        m = same_images & ~band
        assert np.array_equal((fl[m] & NEARCRIT) != 0, minJ[m] < 1e-3), ("NEARCRIT", s, q)
        # SERIES: compare where the margin exceeds the rounding amplified near the critical curve
        d42, d20 = np.abs(B4 - B2), np.abs(B2 - B0)
        with np.errstate(all="ignore"):
            sband = np.abs(d42 / d20 - 1) < 1e-3 * (1 + 1 / np.maximum(minJ, 1e-12) * 1e-6)
        m2 = same_images & ~sband & np.isfinite(d42) & np.isfinite(d20) & (minJ > 1e-6)
        assert np.array_equal((fl[m2] & SERIES) != 0, (d42 > d20)[m2]), ("SERIES", s, q)
        # AMBIG and EXTRA: oracle-side residual of the solver's own roots
        r = np.asarray(roots(s, q, zeta))
        res = oracle.lens_residual(s, q, zeta, r)
        with np.errstate(invalid="ignore", divide="ignore"):
            in_band = ((res > 1e-7) & (res < 1e-5)).any(axis=1)
            edge = (np.abs(np.log10(res) + 7) < 0.01).any(axis=1) | (np.abs(np.log10(res) + 5) < 0.01).any(axis=1)
            near6 = (np.abs(np.log10(res) + 6) < 0.01).any(axis=1)
            plain = (res < 1e-6).sum(axis=1)
        m3 = ~edge
        assert np.array_equal((fl[m3] & AMBIG) != 0, in_band[m3]), ("AMBIG", s, q)
        m4 = ~near6
        extra = (fl >> EXTRA_SHIFT) & 3
        assert np.array_equal(extra[m4], np.minimum(nimg - plain, 3)[m4]), ("EXTRA count", s, q)
        assert np.array_equal((fl[m4] & EXTRA) != 0, (nimg - plain)[m4] > 0), ("EXTRA", s, q)
        stats["n"] += len(zeta)
        stats["nearcrit"] += int(((fl & NEARCRIT) != 0).sum()); stats["series"] += int(((fl & SERIES) != 0).sum())
        stats["ambig"] += int(((fl & AMBIG) != 0).sum()); stats["extra"] += int(((fl & EXTRA) != 0).sum())
        stats["band_excluded"] += int(band.sum() + sband.sum() + edge.sum() + near6.sum())
    report("flag semantics:", stats)
    # every flag was actually exercised, and the excluded bands are a small part of the sample
    assert stats["nearcrit"] > 50 and stats["series"] > 20 and stats["ambig"] > 20 and stats["extra"] > 100, stats
    assert stats["band_excluded"] < 0.03 * stats["n"], stats
    return stats
