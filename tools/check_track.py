"""Dev script: warm-started column strips (host emulation of k_map's control flow) vs cold solves."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import ctypes, sys
import numpy as np
sys.path.insert(0,_ROOT); sys.path.insert(0,_ROOT + '/tools')
import check_emu as ce
libs = ctypes.CDLL(_ROOT + '/tests/host_emu/libmlm_emu_stats.so')
dp=ctypes.POINTER(ctypes.c_double); up=ctypes.POINTER(ctypes.c_ubyte)
def strip(s,q,rho,G,zeta,L):
    zeta=np.ascontiguousarray(zeta,dtype=np.complex128); n=zeta.size
    A=[np.zeros(n) for _ in range(3)]; ni=np.zeros(n,np.uint8); fl=np.zeros(n,np.uint8); nc=ctypes.c_long(0)
    libs.emu_strip(ctypes.c_double(s),ctypes.c_double(q),ctypes.c_double(rho),ctypes.c_double(G),zeta.view(np.float64).ctypes.data_as(dp),ctypes.c_long(n),ctypes.c_int(L),*[a.ctypes.data_as(dp) for a in A],ni.ctypes.data_as(up),fl.ctypes.data_as(up),ctypes.byref(nc))
    return ni,np.stack(A,1),fl,nc.value
def stats(reset=1):
    st=np.zeros(4,np.int64); libs.emu_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_long)),reset); return st
if __name__=="__main__":
  for name,s,q,cx,half,nx in (("C4",1.0,0.5,-0.08,0.8,8192),("C3",0.8,0.1,0.11,0.75,4096)):
    rng=np.random.default_rng(1)
    cols=rng.integers(0,nx,24)
    dy=2*half/nx
    y=-half+(np.arange(nx)+0.5)*dy
    tot_cold=0; tot=0; worst=0; nmis=0
    L=32
    for c in cols:
        x=(cx-half+(c+0.5)*dy)-s*q/(1+q)
        zeta=x+1j*y
        ni,A,fl,nc=strip(s,q,1e-3,0.5,zeta,L)
        ce.lib=libs
        n0,B0,B2,B4,f0=ce.emu_points(s,q,1e-3,0.5,zeta)
        B=np.stack([B0,B2,B4],1)
        nmis+=(ni!=n0).sum()
        rel=np.abs(A/B-1).max(axis=1)
        worst=max(worst,np.nanmax(rel[ni==n0]))
        tot_cold+=nc; tot+=nx
    print(name,'L',L,'cold frac',tot_cold/tot,'(1/L =',1/L,') nimg mismatch vs cold',nmis,'max rel vs cold',worst)
  for name,s,q,cx,half,nx in (("C4",1.0,0.5,-0.08,0.8,8192),):
    dy=2*half/nx; y=-half+(np.arange(nx)+0.5)*dy
    stats()
    tot=0
    for c in range(100,8192,400):
        x=(cx-half+(c+0.5)*dy)-s*q/(1+q)
        strip(s,q,1e-3,0.5,x+1j*y,32); tot+=nx
    st=stats()
    print('tracked: laguerre evals/pt',st[0]/tot,'newton evals/pt',st[1]/tot)
    print('       lens evals/pt', st[2]/tot, 'track fallbacks/pt', st[3]/tot)
