"""GPU-backed drop-in for ``microlensing.multipoles.multipoles`` (reference
``microlensing/multipoles/multipoles.py``): same names, signatures and return values;
every function runs hand-written CUDA through the C ABI of libmlmultipoles.so.

Associated publication: Cassan, A. (2017), MNRAS 468, 3993.
"""
from __future__ import annotations

import numpy as np

from .. import _capi
from ..batched import magnification_points

__all__ = ["quadrupole", "hexadecapole", "_akl", "_solvelenseq", "example"]


def _from_wk(Wk, rho, Gamma, order):
    Wk = np.asarray(Wk)
    nrow = 5 if order == 2 else 7
    if Wk.ndim < 1 or Wk.shape[0] < nrow:
        # the reference indexes Wk[2..4] / Wk[2..6] and fails the same way
        raise IndexError(f"index {nrow - 1} is out of bounds for axis 0 with size {Wk.shape[0] if Wk.ndim else 0}")
    # rows 0-1 are never read (np.empty garbage in the reference, multipoles.py:186); the sums of
    # the reference are axis-less (multipoles.py:62-63,144-146): every trailing axis is an "image".
    W = np.ascontiguousarray(Wk[:nrow], dtype=np.complex128).reshape(nrow, -1)
    ctx = _capi.context()
    out = np.empty(3 if order == 4 else 2, dtype=np.float64)
    ctx.check(ctx.lib.mlm_from_wk_host(ctx.handle, _capi.ptr(W), W.shape[1], order, float(rho), float(Gamma),
                                       _capi.ptr(out)), "mlm_from_wk_host")
    return out


def quadrupole(Wk, rho, Gamma):
    """Quadrupole expansion of finite-source magnification (reference multipoles.py:12-65).

    Parameters
    ----------
    Wk : complex 2D array, shape (7, n) -- derivatives of the lens equation, rows 2..4 are read.
    rho : float -- source radius in Einstein units.
    Gamma : float -- linear limb-darkening coefficient.

    Returns
    -------
    out : float 1D array ``[A0, A2]``.
    """
    return _from_wk(Wk, rho, Gamma, 2)


def hexadecapole(Wk, rho, Gamma):
    """Hexadecapole expansion of finite-source magnification (reference multipoles.py:67-148).

    Same parameters as :func:`quadrupole` (rows 2..6 of ``Wk`` are read); returns
    ``[A0, A2, A4]``; zero images give ``[0., 0., 0.]``.
    """
    return _from_wk(Wk, rho, Gamma, 4)


def _akl(mu, W2, Qkl):
    """a(p-n, n) factors from Q(p-n, n): ``mu*(conj(Qkl) + conj(W2)*Qkl)`` (multipoles.py:150-154)."""
    mu_b, W2_b, Q_b = np.broadcast_arrays(np.asarray(mu), np.asarray(W2), np.asarray(Qkl))
    shape = Q_b.shape
    m = np.ascontiguousarray(mu_b, dtype=np.complex128).ravel()
    w = np.ascontiguousarray(W2_b, dtype=np.complex128).ravel()
    qq = np.ascontiguousarray(Q_b, dtype=np.complex128).ravel()
    out = np.empty(m.size, dtype=np.complex128)
    ctx = _capi.context()
    ctx.check(ctx.lib.mlm_akl_host(ctx.handle, _capi.ptr(m), _capi.ptr(w), _capi.ptr(qq), m.size, _capi.ptr(out)),
              "mlm_akl_host")
    out = out.reshape(shape)
    return out if shape else out[()]


def _solvelenseq(s, q, zeta):
    """Solve the binary-lens equation [convention Cassan (2008)] (multipoles.py:156-165).

    ``zeta`` is ONE complex source position in the primary-lens frame; returns all complex
    roots of the fifth-order lens polynomial (5; 4 when zeta = 0 or zeta = -s exactly, as
    ``np.roots`` does), physical or not, in unspecified order.
    """
    if np.ndim(zeta) != 0:
        raise ValueError("Input must be a rank-1 array.")    # what np.roots raises in the reference
    z = np.array([complex(zeta)], dtype=np.complex128)
    roots = np.empty(5, dtype=np.complex128)
    ctx = _capi.context()
    ctx.check(ctx.lib.mlm_solve_host(ctx.handle, float(s), float(q), _capi.ptr(z), 1, _capi.ptr(roots)),
              "mlm_solve_host")
    return roots[~np.isnan(roots.real)]


def example():
    """Example (reference multipoles.py:167-204): prints A0, A2 and A0, A2, A4 for
    s=1.7, q=0.2, rho=0.01, Gamma=0, source centre -0.5 (centre-of-mass frame)."""
    # binary lens and source parameters
    s, q, rho, Gamma = 1.7, 0.2, 0.01, 0.

    # centre of source: centre of mass -> primary-lens frame (multipoles.py:177-178)
    zeta0 = complex(-.5, 0.) - complex(s * q / (1. + q), 0.)

    # monopole (A0) + quadrupole (A2): fused GPU path, order 2
    r = magnification_points(s, q, rho, Gamma, zeta0, order=2)
    print(" A0, A2 = ", r["A0"][()], r["A2"][()])
    print(" ... should give: 6.08605 6.17629")

    # monopole (A0) + quadrupole (A2) + hexadecapole (A4): fused GPU path, order 4
    r = magnification_points(s, q, rho, Gamma, zeta0, order=4)
    print(" A0, A2, A4 = ", r["A0"][()], r["A2"][()], r["A4"][()])
    print(" ... should give: 6.08605 6.17629 6.18172")


if __name__ == "__main__":
    example()
