"""Offline fuzz of k_vg_partition's row sweep on the HOST (ccb_selftest_vg_partition, no GPU): random and adversarial
site sets (half-integer lattices, a site mirrored across a line through a cell centre, clusters 1e-1 .. 1e-11 wide,
nearly equal py / px for all sites); every unflagged row must equal the stated rule bit for bit.

    python scripts/fuzz_vg_partition.py SEED COUNT      (15 000 sets / 3.9 M rows: no difference, 14 % of the rows flagged)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np
from test_vg_partition_cpu import _check
seed=int(sys.argv[1]); rng=np.random.default_rng(seed)
nflag=rows=0
for it in range(int(sys.argv[2])):
    N=int(rng.integers(2,130))*4
    R=int(rng.integers(1,65))
    res=float(rng.choice([1.0,1.0,1.0,0.5,2.0,0.1,3.0]))
    mode=it%6
    pos=rng.uniform(-0.05*N*res,1.05*N*res,(R,2))
    if mode==1: pos=np.round(pos/res*2)/2*res
    elif mode==2:
        c=(np.floor(rng.uniform(0,N,2))+0.5)*res; th=rng.uniform(0,np.pi); nrm=np.array([np.cos(th),np.sin(th)])
        if R>1: pos[1]=pos[0]-2*np.dot(pos[0]-c,nrm)*nrm+rng.normal(0,10.0**(-int(rng.integers(5,17))),2)
    elif mode==3:
        k=max(1,R//2); pos[:k]=pos[0]+rng.normal(0,10.0**(-int(rng.integers(1,12))),(k,2))
    elif mode==4:
        pos[:,1]=pos[0,1]+rng.normal(0,10.0**(-int(rng.integers(3,16))),R)   # nearly equal py
    elif mode==5:
        pos[:,0]=pos[0,0]+rng.normal(0,10.0**(-int(rng.integers(3,16))),R)   # nearly equal px
    try:
        nflag+=_check(pos,N,res); rows+=N
    except AssertionError as ex:
        np.save(f'fuzz_fail_{seed}_{it}.npy', pos); print('FAIL',seed,it,N,R,res,mode,ex); sys.exit(1)
print('ok',seed,nflag,rows)
