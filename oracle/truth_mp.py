"""40-digit evaluation of the SAME formulae as the reference.  TEST INFRASTRUCTURE ONLY.

Adjudicates parity disputes (SURVEY.md §8(c) rule 4): at small q the reference's
own FP64 noise (``np.roots`` root errors of 1e-12..1e-11, amplified by the
1/z^k, 1/(z+s)^k powers of ``multipoles.py:197-201``) can exceed the 1e-9 parity
bar.  This evaluates multipoles.py:156-165 (polynomial, roots), :182-183
(selection), :197-201 (W_k), :91-146 (a_kl recursion, mu2, mu4, A0/A2/A4) in
mpmath arithmetic, using the partition tables of ``cassan_oracle.q_monomials``
instead of the expanded expressions.
"""
from __future__ import annotations

import mpmath as mp

from .cassan_oracle import q_monomials

_TABLES = {(p - n, n): q_monomials(p - n, n) for p in range(2, 6) for n in range(p + 1)}


def truth_point(s, q, rho, Gamma, zeta, dps=40, return_images=False):
    """(nimg, A0, A2, A4) as Python floats, computed at ``dps`` digits."""
    with mp.workdps(dps):
        s_, q_, rho_, G_ = mp.mpf(s), mp.mpf(q), mp.mpf(rho), mp.mpf(Gamma)
        z_ = mp.mpc(complex(zeta).real, complex(zeta).imag)
        zb = mp.conj(z_)
        m = 1 + q_
        az2 = abs(z_)**2
        # multipoles.py:159-164
        coefs = [
            m**2 * (s_ + zb) * zb,
            m * (s_ * (q_ - az2 * m) + m * ((1 + 2 * s_**2) - az2 + 2 * s_ * zb) * zb),
            m * (s_**2 * q_ - s_ * m * z_ + (2 * s_ + s_**3 * m + s_**2 * m * zb) * zb
                 - 2 * az2 * m * (1 + s_**2 + s_ * zb)),
            -m * (s_ * q_ + s_**2 * (q_ - 1) * zb + (m + s_**2 * (2 + q_)) * z_
                  + az2 * (2 * s_ * (2 + q_) + s_**2 * m * (s_ + zb))),
            -s_ * m * ((2 + s_**2) * z_ + 2 * s_ * az2) - s_**2 * q_,
            -s_**2 * z_,
        ]
        # np.roots strips exact leading zeros (degree drop at zeta=-s, zeta=0) and turns
        # exact trailing zeros into roots at z=0, whose residual is NaN -> rejected (:183)
        while coefs and coefs[0] == 0:
            coefs.pop(0)
        while coefs and coefs[-1] == 0:
            coefs.pop()
        roots = mp.polyroots(coefs, maxsteps=500, extraprec=4 * dps)
        A0 = A2 = A4 = mp.mpf(0)
        n = 0
        imgs = []
        for z in roots:
            if z == 0 or z + s_ == 0:      # pole of W1: NaN/inf residual in FP64, rejected (:183)
                continue
            W1 = (1 / z + q_ / (z + s_)) / m
            if not abs(z - mp.conj(W1) - z_) < mp.mpf("0.000001"):   # :183
                continue
            n += 1
            imgs.append(complex(z))
            W = {}
            fact = mp.mpf(1)
            for k in range(2, 7):                                    # :197-201
                fact *= -(k - 1)
                W[k] = fact / m * (z**(-k) + q_ * (z + s_)**(-k))
            cW2 = mp.conj(W[2])
            mu0 = 1 / (1 - abs(W[2])**2)                             # :91
            a = {(1, 0): mu0 * (1 + cW2), (0, 1): mu0 * mp.mpc(0, 1) * (1 - cW2)}   # :94-95
            for p in range(2, 6):
                Qs = {}
                for nn in range(p + 1):
                    kl = (p - nn, nn)
                    Q = mp.mpc(0)
                    for j, mons in _TABLES[kl].items():
                        inner = mp.mpc(0)
                        for mono, c in mons.items():
                            t = mp.mpf(c)
                            for blk in mono:
                                t = t * a[blk]
                            inner += t
                        Q += W[j + 1] * inner
                    Qs[kl] = Q
                for kl, Q in Qs.items():
                    a[kl] = mu0 * (mp.conj(Q) + cW2 * Q)             # :150-154
            c = mp.conj
            mu2 = mp.im(a[0, 1] * c(a[1, 2] + a[3, 0]) + c(a[1, 0]) * (a[0, 3] + a[2, 1])
                        + 2 * a[0, 2] * c(a[1, 1]) + 2 * a[1, 1] * c(a[2, 0])) / 4      # :142
            mu4 = mp.im((a[0, 5] + 2 * a[2, 3] + a[4, 1]) * c(a[1, 0]) + 4 * a[0, 4] * c(a[1, 1])
                        + 4 * (a[1, 3] + a[3, 1]) * c(a[2, 0]) + 6 * a[1, 2] * c(a[2, 1])
                        + 6 * a[2, 1] * c(a[3, 0]) + a[0, 3] * (6 * c(a[1, 2]) + 2 * c(a[3, 0]))
                        + 4 * a[0, 2] * c(a[1, 3] + a[3, 1]) + 4 * a[1, 1] * c(a[4, 0])
                        + a[0, 1] * c(a[1, 4] + 2 * a[3, 2] + a[5, 0])) / 8             # :143
            t2 = mu0 + mu2 / 2 * (1 - G_ / 5) * rho_**2
            t4 = t2 + mu4 / 24 * (1 - 11 * G_ / 35) * rho_**4
            A0 += abs(mu0)
            A2 += abs(t2)
            A4 += abs(t4)
        out = (n, float(A0), float(A2), float(A4))
        if return_images:
            return out + (imgs,)
        return out
