"""Independent derivation of the operator form used by image_terms_from_W (mlm_core.cuh).

mu0 = 1/(1 - w wbar), w = W2(z) analytic in z.  With the image-plane derivations
    D    = d/dz + w d/dzbar       (d/dzeta    = mu0 D)
    Dbar = wbar d/dz + d/dzbar    (d/dzetabar = mu0 Dbar)
acting on polynomials in (mu, w_k, wbar_k), w_k = W_{k+2}:  D w_k = w_{k+1}, D wbar_k = w wbar_{k+1},
Dbar w_k = wbar w_{k+1}, Dbar wbar_k = wbar_{k+1}, D mu = mu^2 D(w wbar), the source-plane Laplacian is
L = d/dzeta d/dzetabar = mu D (mu Dbar .), and   mu2 = L mu0,   mu4 = 2 L^2 mu0   (DESIGN.md §4).
sympy expands both from these rules alone; the script prints the number of monomials and checks the result
numerically against the formulas coded in the kernel (through the host build) and against the oracle's
literal (xi, eta) recursion of the reference.
"""
import os, sys
import numpy as np
import sympy as sp

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

K = 5
mu = sp.Symbol("mu")
w = sp.symbols("w0:%d" % (K + 1))          # w0 = W2, w1 = W3, ...
wb = sp.symbols("wb0:%d" % (K + 1))


def _deriv(expr, dw, dwb):
    """derivation with given images of the generators; mu follows from mu = 1/(1 - w0 wb0)"""
    out = 0
    for k in range(K):
        out += sp.diff(expr, w[k]) * dw(k) + sp.diff(expr, wb[k]) * dwb(k)
    dn = wb[0] * dw(0) + w[0] * dwb(0)                       # derivative of n = w0 wb0
    out += sp.diff(expr, mu) * mu ** 2 * dn
    return sp.expand(out)


def D(e):
    return _deriv(e, lambda k: w[k + 1], lambda k: w[0] * wb[k + 1])


def Dbar(e):
    return _deriv(e, lambda k: wb[0] * w[k + 1], lambda k: wb[k + 1])


def L(e):
    return sp.expand(mu * D(mu * Dbar(e)))


def derive():
    mu2 = L(mu)
    mu4 = 2 * L(mu2)
    return mu2, mu4


def evaluate(expr, W):
    """W: dict k -> complex arrays for W2..W6"""
    subs = [mu] + list(w[:5]) + list(wb[:5])
    f = sp.lambdify(subs, expr, "numpy")
    m0 = 1.0 / (1.0 - np.abs(W[2]) ** 2)
    args = [m0] + [W[k + 2] for k in range(5)] + [np.conj(W[k + 2]) for k in range(5)]
    return f(*args)


if __name__ == "__main__":
    from oracle import cassan_oracle as o
    mu2, mu4 = derive()
    print("monomials: mu2", len(sp.Add.make_args(mu2)), " mu4", len(sp.Add.make_args(mu4)))
    rng = np.random.default_rng(0)
    z = rng.normal(size=100) + 1j * rng.normal(size=100)
    Wk = o.wk_table(1.0, 0.5, z, 6)
    W = {k: Wk[k] for k in range(2, 7)}
    r0, r2, r4 = o._per_image(Wk, 5)
    e2, e4 = evaluate(mu2, W), evaluate(mu4, W)
    print("imaginary parts (must vanish):", np.abs(e2.imag / e2.real).max(), np.abs(e4.imag / e4.real).max())
    print("vs reference recursion (oracle): mu2", np.abs(e2.real / r2 - 1).max(), " mu4", np.abs(e4.real / r4 - 1).max())
