"""Prototype: closed-form Laplacian / bi-Laplacian of the image magnification (operator form) vs the
oracle's literal (xi, eta) recursion.  D = d_z + w d_zbar, Dbar = wbar d_z + d_zbar, d_zeta = mu D."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import sys
import numpy as np
sys.path.insert(0, _ROOT)
from oracle import cassan_oracle as o


def closed(W):
    w, w1, w2, w3, w4 = W[2], W[3], W[4], W[5], W[6]
    c = np.conj
    n = np.abs(w) ** 2
    mu = 1. / (1. - n)
    p = np.abs(w1) ** 2
    wb2 = c(w) ** 2
    wb3 = wb2 * c(w)
    A = wb2 * w1 + w * c(w1)
    B = p * (1 + 2 * n) + 2 * np.real(wb2 * w2)
    C = 3 * c(w) * p + wb3 * w2 + w * c(w2)
    U = c(w1) * w2
    X = c(w) * U
    Y = w1 * c(w2)
    Z = wb3 * w3 + c(wb2) * c(w3)
    a2 = np.abs(A) ** 2
    ww1b = w * c(w1)
    S = (45 + 33 * n) * X + (15 + 57 * n) * Y + 24 * p * A + 9 * p * ww1b + 15 * Z      # 12 E + 3 DC
    T = np.real(c(w) * c(w1) * w3)
    Dp = U + w * Y
    ReDE = (np.real(c(A) * (2 * X + 4 * Y)) + (3 + 2 * n) * np.real(ww1b * U) + (1 + 7 * n + 2 * n * n) * np.abs(w2) ** 2
            + (6 + 9 * n) * T + 2 * np.real(A * Dp) + 2 * p * B + 2 * np.real(wb3 * w4))
    lap1 = 3 * mu * a2 + B
    lap2 = mu * (mu * (mu * 105 * a2 * a2 + 72 * a2 * B + 33 * np.real(c(A) ** 2 * C)) + 7 * B * B + 3 * np.abs(C) ** 2
                 + np.real(c(A) * S)) + ReDE
    return mu, mu ** 4 * lap1, 2 * mu ** 6 * lap2


rng = np.random.default_rng(0)
for s, q in ((1.0, 0.5), (0.8, 0.1), (1.0, 1e-3)):
    z = rng.normal(size=200) + 1j * rng.normal(size=200)
    Wk = o.wk_table(s, q, z, 6)
    mu0, mu2, mu4 = o._per_image(Wk, 5)
    m0, m2, m4 = closed(Wk)
    print(s, q, np.abs(m0 / mu0 - 1).max(), np.abs(m2 / mu2 - 1).max(), np.abs(m4 / mu4 - 1).max(),
          np.median(np.abs(m4 / mu4 - 1)))
