"""Dev script: large random comparison of the host-emulated kernel logic vs batch oracle."""
import os as _os
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
import ctypes, sys, time
import numpy as np
sys.path.insert(0,_ROOT); sys.path.insert(0,_ROOT + '/tools')
from oracle import cassan_oracle as o
import check_emu as ce
libs = ctypes.CDLL(_ROOT + '/tests/host_emu/libmlm_emu_stats.so')
def run(name,s,q,rho,G,zeta):
    ce.lib = libs
    st=np.zeros(4,np.int64); libs.emu_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_long)),1)
    t=time.time(); ni,A0,A2,A4,fl=ce.emu_points(s,q,rho,G,zeta); te=time.time()-t
    libs.emu_stats(st.ctypes.data_as(ctypes.POINTER(ctypes.c_long)),1)
    nb,B0,B2,B4,z,keep,minJ,res=o.batch_a4(s,q,rho,G,zeta,return_extra=True)
    bad=ni!=nb
    amb=((res>1e-7)&(res<1e-5)).any(axis=1)
    with np.errstate(all='ignore'):
        rel=np.maximum(np.abs(A0/B0-1),np.maximum(np.abs(A2/B2-1),np.abs(A4/B4-1)))
    ok=~bad
    print(f"{name}: N={zeta.size} f5={np.mean(nb==5):.3f} nimg mismatch={bad.sum()} (ambiguous {np.sum(bad&amb)}) "
          f"rel>1e-9: {np.sum(ok&(rel>1e-9))} max rel {np.nanmax(rel[ok]):.2e} flags noconv={np.sum(fl&1>0)} viete={np.sum(fl&2>0)} "
          f"lag/pt={st[0]/zeta.size:.2f} newt/pt={st[1]/zeta.size:.2f} emu {te/zeta.size*1e6:.2f} us/pt")
    return ni,nb,rel,fl
rng=np.random.default_rng(5)
N=200000
# C4
s,q=1.0,0.5
z=(-0.08+rng.uniform(-.8,.8,N))+1j*rng.uniform(-.8,.8,N)-s*q/(1+q)
run('C4',s,q,1e-3,0.5,z)
s,q=0.8,0.1
z=(0.11+rng.uniform(-.75,.75,N))+1j*rng.uniform(-.75,.75,N)-s*q/(1+q)
run('C3',s,q,1e-3,0.5,z)
s,q=1.0,1e-3
tau=rng.uniform(-2,2,N); z=(tau+0.01j)*np.exp(0.35j)-s*q/(1+q)
run('C2',s,q,1e-3,0.5,z)
z=rng.uniform(-.05,.05,N)+1j*rng.uniform(-.05,.05,N)-s*q/(1+q)
run('C2core',s,q,1e-3,0.5,z)
for k in range(12):
    s=float(np.exp(rng.uniform(np.log(0.5),np.log(2)))); q=float(np.exp(rng.uniform(np.log(1e-4),0)))
    z=rng.uniform(-2,2,20000)+1j*rng.uniform(-2,2,20000)
    run(f'rand s={s:.3f} q={q:.2e}',s,q,1e-3,0.5,z)

def adjudicate(name,s,q,rho,G,zeta,nmax=60):
    from oracle.truth_mp import truth_point
    ni,nb,rel,fl=run(name,s,q,rho,G,zeta)
    A=ce.emu_points(s,q,rho,G,zeta); B=o.batch_a4(s,q,rho,G,zeta,return_extra=True)
    idx=np.where((rel>1e-9)|(ni!=nb))[0]
    idx=idx[np.argsort(-np.nan_to_num(rel[idx],nan=1))][:nmax]
    worse=0; cntbad=0
    for i in idx:
        t=truth_point(s,q,rho,G,zeta[i])
        eg=max(abs(A[k][i]/t[k]-1) for k in (1,2,3)); er=max(abs(B[k][i]/t[k]-1) for k in (1,2,3))
        if A[0][i]!=t[0]: cntbad+=1
        if eg>max(er,1e-9): worse+=1; print('  WORSE', zeta[i], A[0][i],B[0][i],t[0], eg, er, B[6][i], fl[i])
    print(f'  adjudicated {len(idx)}: gpu-logic worse than ref in {worse}, nimg!=truth in {cntbad}')
if __name__=="__main__":
    s,q=1.0,0.5
    z=(-0.08+rng.uniform(-.8,.8,N))+1j*rng.uniform(-.8,.8,N)-s*q/(1+q)
    adjudicate('C4',s,q,1e-3,0.5,z)
    s,q=1.0,1e-3
    z=rng.uniform(-.05,.05,N)+1j*rng.uniform(-.05,.05,N)-s*q/(1+q)
    adjudicate('C2core',s,q,1e-3,0.5,z)
    s,q=1.182,1.69e-4
    z=rng.uniform(-2,2,20000)+1j*rng.uniform(-2,2,20000)
    adjudicate('smallq',s,q,1e-3,0.5,z)
