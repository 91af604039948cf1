"""Dev script: host-emulated kernel logic vs golden fixtures (reference) and truth."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import ctypes, sys, os
import numpy as np
sys.path.insert(0, _ROOT)
from oracle import cassan_oracle as o
lib = ctypes.CDLL(_ROOT + '/tests/host_emu/libmlm_emu.so')
dp = ctypes.POINTER(ctypes.c_double); up = ctypes.POINTER(ctypes.c_ubyte)
def P(a): return a.ctypes.data_as(dp)
def U(a): return a.ctypes.data_as(up)
def emu_points(s,q,rho,G,zeta,order=4):
    zeta=np.ascontiguousarray(zeta,dtype=np.complex128); n=zeta.size
    A0=np.zeros(n);A2=np.zeros(n);A4=np.zeros(n);ni=np.zeros(n,np.uint8);fl=np.zeros(n,np.uint8)
    lib.emu_points(ctypes.c_double(s),ctypes.c_double(q),ctypes.c_double(rho),ctypes.c_double(G),P(zeta.view(np.float64)),ctypes.c_long(n),ctypes.c_int(order),P(A0),P(A2),P(A4),U(ni),U(fl))
    return ni,A0,A2,A4,fl
def emu_roots(s,q,zeta):
    zeta=np.ascontiguousarray(zeta,dtype=np.complex128); n=zeta.size
    r=np.zeros((n,5),np.complex128); fl=np.zeros(n,np.uint8)
    lib.emu_roots(ctypes.c_double(s),ctypes.c_double(q),P(zeta.view(np.float64)),ctypes.c_long(n),P(r.view(np.float64)),U(fl))
    return r,fl
if __name__=="__main__":
    # image terms
    rng=np.random.default_rng(0)
    z=rng.normal(size=200)+1j*rng.normal(size=200)
    Wk=o.wk_table(1.0,0.5,z,6)
    mu0,mu2,mu4=o._per_image(Wk,5)
    W=np.ascontiguousarray(Wk[2:7].T)
    for lit in (0,1):
        out=np.zeros((200,3))
        lib.emu_image_terms(P(W.view(np.float64)),ctypes.c_long(200),ctypes.c_int(lit),P(out))
        print('literal' if lit else 'wirtinger', np.abs(out[:,0]/mu0-1).max(), np.abs(out[:,1]/mu2-1).max(), np.abs(out[:,2]/mu4-1).max())
    g=np.load(_ROOT + '/tests/golden/golden_points.npz')
    tags=g['tags']; par=g['params']; zeta=g['zeta']
    res=np.zeros((len(tags),3)); ni=np.zeros(len(tags),int); fl=np.zeros(len(tags),int)
    # group by params
    keys={}
    for i,p in enumerate(map(tuple,par)): keys.setdefault(p,[]).append(i)
    for p,idx in keys.items():
        idx=np.array(idx)
        n_,A0,A2,A4,f_=emu_points(*p,zeta[idx])
        res[idx,0]=A0;res[idx,1]=A2;res[idx,2]=A4;ni[idx]=n_;fl[idx]=f_
    H=g['hexadecapole'];T=g['truth']
    for t in np.unique(tags):
        m=tags==t
        with np.errstate(all='ignore'):
            rr=np.abs(res[m]/H[m]-1); rt=np.abs(res[m]/T[m]-1)
        print(t, m.sum(), 'nimg!=ref', (ni[m]!=g['nimg'][m]).sum(), 'nimg!=truth', (ni[m]!=g['truth_nimg'][m]).sum(),
              'max rel vs ref', np.nanmax(rr,axis=0), 'vs truth', np.nanmax(rt,axis=0), 'flags', np.bincount(fl[m],minlength=1)[:64].nonzero()[0], (fl[m]&3).astype(bool).sum())
    m=tags=='edge'
    print(ni[m], g['nimg'][m], res[m][:, 0], H[m][:,0], fl[m])
