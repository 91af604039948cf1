"""CPU oracle for the Cassan (2017) binary-lens A0/A2/A4 path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference algorithm in
``/root/reference/microlensing/multipoles/multipoles.py`` (cited below as
``multipoles.py:LINE``).  It exists to *check* the CUDA path; it is never
imported by the product package ``microlensing_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.

Pinning: the reference ships no tests (``.travis.yml:7`` "No test yet"); the
only pinned numbers are the 6-digit "should give" strings of ``example()``
(``multipoles.py:194,204``).  The oracle is therefore pinned against outputs
of the LIVE reference run in the build container (``tests/golden/make_golden.py``
imports ``/root/reference`` and writes ``tests/golden/*.npz``); the committed
fixtures travel to the GPU box, the reference does not.
``tests/test_oracle_golden.py`` checks this module against those fixtures.

The root solve of the reference is ``numpy.roots`` (``multipoles.py:165``), i.e.
the third-party NumPy (unpinned by the reference, ``setup.py:59``; 2.3.5 with
OpenBLAS 0.3.30 here): strip leading/trailing zero coefficients, build the
companion matrix, ``numpy.linalg.eigvals`` (LAPACK ``zgeev``).  The scalar
entry points below call ``np.roots`` exactly as the reference does; the
batched entry point builds the same companion matrices and calls the same
``eigvals`` gufunc, so its roots are the roots ``np.roots`` would return.

How the Q factors are obtained: instead of carrying the expanded formulas of
``multipoles.py:98-133`` the oracle regenerates them.  ``Q_{k,l}`` (k xi- and l
eta-derivatives of conj(W1(z(zeta)))) is a Faa di Bruno sum over the set
partitions of the p=k+l derivative labels into j>=2 blocks,
``Q_kl = sum_j W_{j+1} * sum_partitions prod_blocks a_{block}``
(the single-block term is the ``W2 * a_kl`` that ``_akl`` solves for,
``multipoles.py:150-154``).  Identical monomials are collected, which yields the
integer coefficients visible in the reference (3, 6, 10, 15, ...).  Python
source for the collected forms is emitted once at import and ``exec``-ed, so
the per-position execution profile (a few hundred tiny NumPy ufunc calls on
length-3/5 arrays) is the reference's.
"""
from __future__ import annotations

import itertools
from collections import Counter

import numpy as np

__all__ = [
    "lens_polynomial", "solvelenseq", "akl", "quadrupole", "hexadecapole",
    "select_images", "wk_table", "point_a4", "point_a2", "batch_a4",
    "example_values", "q_monomials", "cm_to_primary",
]

RESIDUAL_THRESHOLD = 0.000001  # multipoles.py:183 (absolute)


# ----------------------------------------------------------------------------
# Q-factor generator (replaces the pasted expansions multipoles.py:98-133; the
# reference obtained them from Rkp.py:37-47,75-97 by string rewriting)
# ----------------------------------------------------------------------------
def _set_partitions(items):
    """All set partitions of a list of distinct labels."""
    if len(items) == 1:
        yield [items]
        return
    first, rest = items[0], items[1:]
    for part in _set_partitions(rest):
        for i in range(len(part)):
            yield part[:i] + [[first] + part[i]] + part[i + 1:]
        yield [[first]] + part


def q_monomials(k: int, l: int):
    """Collected expansion of Q_{k,l}: {j: Counter{((k1,l1),(k2,l2),..): coeff}}.

    j is the number of blocks; the factor multiplying the block sum is W_{j+1}.
    """
    labels = ["x"] * k + ["e"] * l
    idx = list(range(k + l))
    out: dict[int, Counter] = {}
    for part in _set_partitions(idx):
        j = len(part)
        if j < 2:
            continue
        mono = tuple(sorted(
            (sum(labels[i] == "x" for i in blk), sum(labels[i] == "e" for i in blk))
            for blk in part))
        out.setdefault(j, Counter())[mono] += 1
    return out


def _emit_source(pmax: int) -> str:
    """Python source of a function computing the per-image mu0, mu2[, mu4]."""
    lines = ["def _per_image(W, pmax):",
             "    W2 = W[2]",
             "    mu0 = 1. / (1. - np.abs(W2)**2)",                 # multipoles.py:36,91
             "    a10 = mu0 * (1. + np.conj(W2))",                  # :39,94
             "    a01 = mu0 * 1j * (1. - np.conj(W2))",             # :40,95
             ]
    for p in range(2, pmax + 1):
        if p == 4:
            lines.append("    if pmax >= 4:")
        ind = "        " if p >= 4 else "    "
        for n in range(p + 1):
            k, l = p - n, n
            terms = []
            for j, mons in sorted(q_monomials(k, l).items()):
                inner = " + ".join(
                    (f"{c}.*" if c != 1 else "") + "*".join(f"a{a}{b}" for a, b in mono)
                    for mono, c in sorted(mons.items(), reverse=True))
                terms.append(f"W[{j + 1}] * ({inner})")
            lines.append(f"{ind}Q{k}{l} = " + " + ".join(terms))
        for n in range(p + 1):
            k, l = p - n, n
            # multipoles.py:150-154
            lines.append(f"{ind}a{k}{l} = mu0 * (np.conj(Q{k}{l}) + np.conj(W2) * Q{k}{l})")
    # multipoles.py:61,142
    lines.append("    mu2 = 1./4. * np.imag(a01 * np.conj(a12 + a30) + np.conj(a10) * (a03 + a21)"
                 " + 2. * a02 * np.conj(a11) + 2. * a11 * np.conj(a20))")
    lines.append("    if pmax < 4:")
    lines.append("        return mu0, mu2, None")
    # multipoles.py:143
    lines.append("    mu4 = 1./8. * np.imag((a05 + 2. * a23 + a41) * np.conj(a10) + 4. * a04 * np.conj(a11)"
                 " + 4 * (a13 + a31) * np.conj(a20) + 6. * a12 * np.conj(a21) + 6. * a21 * np.conj(a30)"
                 " + a03 * (6. * np.conj(a12) + 2. * np.conj(a30)) + 4. * a02 * np.conj(a13 + a31)"
                 " + 4. * a11 * np.conj(a40) + a01 * np.conj(a14 + 2. * a32 + a50))")
    lines.append("    return mu0, mu2, mu4")
    return "\n".join(lines)


_SOURCE = _emit_source(5)
_ns: dict = {"np": np}
exec(compile(_SOURCE, "<cassan_oracle generated>", "exec"), _ns)
_per_image = _ns["_per_image"]


# ----------------------------------------------------------------------------
# The five reference functions, restated
# ----------------------------------------------------------------------------
def akl(mu, W2, Qkl):
    """multipoles.py:150-154."""
    return mu * (np.conj(Qkl) + np.conj(W2) * Qkl)


def quadrupole(Wk, rho, Gamma):
    """[A0, A2]; multipoles.py:12-65 (reads Wk[2..4], axis-less sum :62-63)."""
    mu0, mu2, _ = _per_image(Wk, 3)
    A0 = np.sum(np.abs(mu0))
    A2 = np.sum(np.abs(mu0 + 1. / 2. * mu2 * (1. - 1. / 5. * Gamma) * rho**2))
    return np.array([A0, A2])


def hexadecapole(Wk, rho, Gamma):
    """[A0, A2, A4]; multipoles.py:67-148 (reads Wk[2..6], sums :144-146)."""
    mu0, mu2, mu4 = _per_image(Wk, 5)
    A0 = np.sum(np.abs(mu0))
    A2 = np.sum(np.abs(mu0 + 1. / 2. * mu2 * (1. - 1. / 5. * Gamma) * rho**2))
    A4 = np.sum(np.abs(mu0 + 1. / 2. * mu2 * (1. - 1. / 5. * Gamma) * rho**2
                       + 1. / 24. * mu4 * (1. - 11. / 35. * Gamma) * rho**4))
    return np.array([A0, A2, A4])


def lens_polynomial(s, q, zeta):
    """Six coefficients, highest degree first; multipoles.py:159-164.

    Works for scalar or array ``zeta`` (returns shape (6,) + zeta.shape).
    """
    zeta = np.asarray(zeta, dtype=np.complex128)
    zb = np.conj(zeta)
    m = 1. + q
    az2 = np.abs(zeta)**2
    c5 = m**2 * (s + zb) * zb
    c4 = m * (s * (q - az2 * m) + m * ((1. + 2 * s**2) - az2 + 2. * s * zb) * zb)
    c3 = m * (s**2 * q - s * m * zeta + (2 * s + s**3 * m + s**2 * m * zb) * zb
              - 2. * az2 * m * (1. + s**2 + s * zb))
    c2 = - m * (s * q + s**2 * (q - 1.) * zb + (m + s**2 * (2. + q)) * zeta
                + az2 * (2. * s * (2. + q) + s**2 * m * (s + zb)))
    c1 = - s * m * ((2. + s**2) * zeta + 2. * s * az2) - s**2 * q
    c0 = - s**2 * zeta
    return np.array([c5, c4, c3, c2, c1, c0 + 0 * c5])


def solvelenseq(s, q, zeta):
    """All complex roots via numpy.roots, scalar zeta only; multipoles.py:156-165."""
    zeta = np.complex128(zeta)
    return np.roots(list(lens_polynomial(s, q, zeta)))


def cm_to_primary(s, q, zeta_cm):
    """Centre-of-mass source coordinate -> primary-lens frame; multipoles.py:177-178."""
    return zeta_cm - s * q / (1. + q)


def select_images(s, q, zeta, z):
    """Physical images by the absolute residual test; multipoles.py:182-183."""
    W1 = 1. / (1. + q) * (1. / z + q / (z + s))
    return z[np.abs(z - W1.conjugate() - zeta) < RESIDUAL_THRESHOLD]


def lens_residual(s, q, zeta, z):
    """|z - conj(W1(z)) - zeta| of candidate roots z (last axis) at positions zeta: the quantity the reference
    compares with 1e-6; multipoles.py:182-183.  NaN where z sits on a lens, like the reference's 1/0."""
    z = np.asarray(z, dtype=np.complex128)
    zeta = np.asarray(zeta, dtype=np.complex128)
    with np.errstate(all="ignore"):
        W1 = 1. / (1. + q) * (1. / z + q / (z + s))
        return np.abs(z - W1.conjugate() - zeta[..., None])


def wk_table(s, q, z0, kmax=6):
    """Wk[2..kmax] rows of a (7, n) table; multipoles.py:185-191,197-201.

    Rows 0-1 are left as zeros (the reference leaves them uninitialised).
    """
    Wk = np.zeros((7,) + np.shape(z0), dtype=np.complex128)
    fact = 1.
    for k in range(2, kmax + 1):
        fact *= -(k - 1.)            # -1, 2, -6, 24, -120
        Wk[k] = fact / (1. + q) * (1. / z0**k + q / (z0 + s)**k)
    return Wk


def point_a4(s, q, rho, Gamma, zeta):
    """Literal per-position replay of multipoles.py:181-202 -> (nimg, A0, A2, A4)."""
    with np.errstate(all="ignore"):
        z0 = solvelenseq(s, q, zeta)
        z0 = select_images(s, q, zeta, z0)
        Wk = wk_table(s, q, z0, 6)
        A0, A2, A4 = hexadecapole(Wk, rho, Gamma)
    return len(z0), A0, A2, A4


def point_a2(s, q, rho, Gamma, zeta):
    """Per-position replay of multipoles.py:181-192 -> (nimg, A0, A2)."""
    with np.errstate(all="ignore"):
        z0 = solvelenseq(s, q, zeta)
        z0 = select_images(s, q, zeta, z0)
        Wk = wk_table(s, q, z0, 4)
        A0, A2 = quadrupole(Wk, rho, Gamma)
    return len(z0), A0, A2


def batch_roots(s, q, zeta):
    """Roots for an array of zeta with the arithmetic of np.roots (companion + eigvals).

    Requires non-zero leading and trailing coefficients (zeta != 0, zeta != -s),
    i.e. the cases where np.roots strips nothing; other points must go through
    ``solvelenseq``.
    """
    zeta = np.asarray(zeta, dtype=np.complex128).ravel()
    c = lens_polynomial(s, q, zeta)                       # (6, N)
    if np.any(c[0] == 0) or np.any(c[5] == 0):
        raise ValueError("degenerate polynomial: use the scalar path")
    n = zeta.size
    A = np.zeros((n, 5, 5), dtype=np.complex128)
    A[:, 1, 0] = A[:, 2, 1] = A[:, 3, 2] = A[:, 4, 3] = 1.
    A[:, 0, :] = (-c[1:] / c[0]).T                        # numpy roots(): A[0,:] = -p[1:]/p[0]
    return np.linalg.eigvals(A)                           # (N, 5)


def batch_a4(s, q, rho, Gamma, zeta, return_extra=False):
    """Vectorised multipoles.py:181-202 over an array of positions.

    Returns (nimg uint8, A0, A2, A4) arrays of zeta's shape.  Non-selected roots
    contribute exact zeros to the sums, so the values equal the per-position
    replay (same summation order).  With ``return_extra`` also returns
    (roots, keep mask, minJ) for parity masks.
    """
    shape = np.shape(zeta)
    zeta = np.asarray(zeta, dtype=np.complex128).ravel()
    with np.errstate(all="ignore"):
        z = batch_roots(s, q, zeta)                       # (N,5)
        W1 = 1. / (1. + q) * (1. / z + q / (z + s))
        res = np.abs(z - W1.conjugate() - zeta[:, None])
        keep = res < RESIDUAL_THRESHOLD
        Wk = wk_table(s, q, z, 6)                         # (7,N,5)
        mu0, mu2, mu4 = _per_image(Wk, 5)
        t0 = np.abs(mu0)
        t2 = np.abs(mu0 + 1. / 2. * mu2 * (1. - 1. / 5. * Gamma) * rho**2)
        t4 = np.abs(mu0 + 1. / 2. * mu2 * (1. - 1. / 5. * Gamma) * rho**2
                    + 1. / 24. * mu4 * (1. - 11. / 35. * Gamma) * rho**4)
        A0 = _ordered_masked_sum(t0, keep)
        A2 = _ordered_masked_sum(t2, keep)
        A4 = _ordered_masked_sum(t4, keep)
    nimg = keep.sum(axis=1).astype(np.uint8)
    out = (nimg.reshape(shape), A0.reshape(shape), A2.reshape(shape), A4.reshape(shape))
    if return_extra:
        J = np.where(keep, np.abs(1. - np.abs(Wk[2])**2), np.inf)
        return out + (z, keep, J.min(axis=1).reshape(shape), res)
    return out


def _ordered_masked_sum(t, keep):
    """Left-to-right sum over the kept columns (np.sum on <8 elements is sequential)."""
    acc = np.zeros(t.shape[0])
    for j in range(t.shape[1]):
        acc = acc + np.where(keep[:, j], t[:, j], 0.)
    return acc


def example_values():
    """Numbers printed by example(); multipoles.py:171-204."""
    s, q, rho, Gamma = 1.7, 0.2, 0.01, 0.
    zeta0 = cm_to_primary(s, q, complex(-.5, 0.))
    n, A0, A2 = point_a2(s, q, rho, Gamma, zeta0)
    n4, B0, B2, B4 = point_a4(s, q, rho, Gamma, zeta0)
    return (A0, A2), (B0, B2, B4)
