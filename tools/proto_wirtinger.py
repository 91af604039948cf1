"""Prototype: Wirtinger-basis recursion for mu0, mu2, mu4 vs the oracle's (xi,eta) recursion."""
import sys, itertools
import numpy as np
sys.path.insert(0, '/root/repo')
from oracle import cassan_oracle as o
from math import comb

def wirt(W):
    # W[2..6] arrays
    w = W[2]; F = {k: np.conj(W[k+1]) for k in range(1, 6)}
    mu0 = 1./(1.-np.abs(w)**2)
    u = {(0,1): np.ones_like(w), (1,0): w}
    b = {(1,0): np.ones_like(w), (0,1): F[1]}
    need = {2:[(2,0),(1,1),(0,2)], 3:[(3,0),(2,1),(1,2),(0,3)], 4:[(3,1),(2,2),(1,3)], 5:[(3,2),(2,3)]}
    for p in range(2,6):
        R = {}
        for (m,n) in need[p]:
            acc = 0
            for j, mons in o.q_monomials(m,n).items():
                inner = 0
                for mono, c in mons.items():
                    t = c
                    for blk in mono:
                        t = t*u[blk]
                    inner = inner + t
                acc = acc + F[j]*inner
            R[(m,n)] = acc
        for (m,n) in need[p]:
            b[(m,n)] = mu0*(R[(m,n)] + F[1]*np.conj(R[(n,m)]))
        for (m,n) in need[p]:
            u[(m,n)] = np.conj(b[(n,m)])
    # mu2 = d_z d_zb (|b10|^2 - |b01|^2)
    lap1 = 2*np.real(b[2,1]*np.conj(b[1,0])) - 2*np.real(b[1,2]*np.conj(b[0,1])) + np.abs(b[2,0])**2 - np.abs(b[0,2])**2
    lap2 = 0
    for i in range(3):
        for j in range(3):
            cc = comb(2,i)*comb(2,j)
            lap2 = lap2 + cc*( b[1+i,j]*np.conj(b[3-j,2-i]) - b[i,1+j]*np.conj(b[2-j,3-i]) )
    return mu0, mu0**4*lap1, mu0**6*np.real(lap2), np.abs(np.imag(lap2)).max()

rng = np.random.default_rng(0)
s,q=1.0,0.5
z = rng.normal(size=50)+1j*rng.normal(size=50)
Wk = o.wk_table(s,q,z,6)
mu0,mu2,mu4 = o._per_image(Wk,5)
m0,l1,l2,im = wirt(Wk)
print(np.abs(m0-mu0).max(), np.abs(l1/mu2).min(), np.abs(l1/mu2).max(), (l2/mu4).min(), (l2/mu4).max(), im)
print(np.abs(l1/mu2-1).max(), np.abs(2*l2/mu4-1).max())
