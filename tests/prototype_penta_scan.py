#!/usr/bin/env python
"""Prototype (numpy, CPU) for the next Poisson y kernel: chunk-parallel pentadiagonal substitution.

Today one thread marches along the 512 rows of a (kx, kz) mode five times and keeps the intermediate vector and the LU factor
lines in global memory: 37 doubles per (mode, row), 148 B per real point.  The substitution stages
PENTADSS (src/utils/linear5.f90:76-131) are second-order linear recurrences,
    forward   y_n = f_n + b_n y_{n-1} + a_n y_{n-2},        backward  x_n = (y_n + d_n x_{n+1} + e_n x_{n+2}) c_n,
so a chunk of 16 rows maps its two inflow values to its two outflow values by an affine 2x2 map, and the true inflow of every
chunk follows from a prefix scan of those maps over the 32 chunks of a line: one warp per mode, one lane per chunk, five
shuffle steps, the whole line in registers.  Unlike the tridiagonal compact schemes of the line kernels the maps do NOT
decay for small wavenumbers (u' + lambda u = f is an integral: the solution depends on all previous rows), so the scan
cannot be truncated to a window -- this script measures that, and checks that the scan is nevertheless accurate: it
solves the reference's own factored systems (oracle.integral.int1_initialize on the tanh grid of C3) for the range of
eigenvalues of C3 and both boundary types, and compares with PENTADSS.

    python tests/prototype_penta_scan.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import grid_tanh
from oracle import fdm, integral
ny=512
y=grid_tanh(ny)
g=fdm.Plan(y,False,False,name='y')
C=16
def chunked_pentadss(a,b,c,d,e,f):
    n=len(a); T=(n+C-1)//C
    # forward: y_n = f_n + b_n y_{n-1} + a_n y_{n-2}; state s=(y_{n-1}, y_{n-2})
    yl=np.zeros_like(f); Ms=[]; vs=[]
    bounds=[(t*C,min((t+1)*C,n)) for t in range(T)]
    # local zero-inflow + homogeneous responses
    H=np.zeros((n,2))   # response of y_n to inflow (y_-1, y_-2) of its chunk
    for (s0,s1) in bounds:
        p1=0.0;p2=0.0; h1=np.array([1.0,0.0]); h2=np.array([0.0,1.0])  # h1 = dy_{n-1}/d inflow, h2 = dy_{n-2}/d inflow
        for i in range(s0,s1):
            bi = b[i] if i>=1 else 0.0; ai = a[i] if i>=2 else 0.0
            yi=f[i]+bi*p1+ai*p2
            hi=bi*h1+ai*h2
            yl[i]=yi; H[i]=hi
            p2=p1;p1=yi; h2=h1;h1=hi
        Ms.append(np.array([h1,h2])); vs.append((p1,p2))
    # scan
    s=(np.zeros_like(f[0]),np.zeros_like(f[0]))
    yt=np.zeros_like(f)
    normM=max(np.abs(M).max() for M in Ms)
    for t,(s0,s1) in enumerate(bounds):
        for i in range(s0,s1):
            yt[i]=yl[i]+H[i,0]*s[0]+H[i,1]*s[1]
        M=Ms[t]; v=vs[t]
        s=(v[0]+M[0,0]*s[0]+M[0,1]*s[1], v[1]+M[1,0]*s[0]+M[1,1]*s[1])
    # backward: x_n = (y_n + d_n x_{n+1} + e_n x_{n+2}) c_n ; state (x_{n+1}, x_{n+2})
    xl=np.zeros_like(f); G=np.zeros((n,2)); Mb=[];vb=[]
    for (s0,s1) in bounds:
        p1=0.0;p2=0.0;h1=np.array([1.0,0.0]);h2=np.array([0.0,1.0])
        for i in range(s1-1,s0-1,-1):
            di = d[i] if i<=n-2 else 0.0; ei = e[i] if i<=n-3 else 0.0
            xi=(yt[i]+di*p1+ei*p2)*c[i]
            hi=(di*h1+ei*h2)*c[i]
            xl[i]=xi;G[i]=hi
            p2=p1;p1=xi;h2=h1;h1=hi
        Mb.append(np.array([h1,h2]));vb.append((p1,p2))
    normB=max(np.abs(M).max() for M in Mb)
    s=(np.zeros_like(f[0]),np.zeros_like(f[0]))
    x=np.zeros_like(f)
    for t in range(T-1,-1,-1):
        s0,s1=bounds[t]
        for i in range(s0,s1):
            x[i]=xl[i]+G[i,0]*s[0]+G[i,1]*s[1]
        M=Mb[t];v=vb[t]
        s=(v[0]+M[0,0]*s[0]+M[0,1]*s[1], v[1]+M[1,0]*s[0]+M[1,1]*s[1])
    return x,normM,normB
rng=np.random.default_rng(0)
for ibc in (fdm.BCS_MIN, fdm.BCS_MAX):
  for lam in (0.0, 1e-3, 1.0, 10.0, 100.0, 1e3, 1e4, 3e5):
    sgn = 1.0 if ibc==fdm.BCS_MIN else -1.0
    fdmi=integral.int1_initialize(g.der1, sgn*np.sqrt(lam) if lam>0 else 0.0, ibc)
    nx=fdmi.lhs.shape[0]-1
    cols=[np.array(fdmi.lhs[2:nx,k]).reshape(nx-2) for k in range(1,6)]
    f=rng.standard_normal((nx-2,3))
    ref=f.copy(); fdm.pentadss(*cols, ref)
    got,nm,nb=chunked_pentadss(*cols,f)
    print('bc',ibc,'lam',lam,'err',np.linalg.norm(got-ref)/np.linalg.norm(ref),'|Mfwd|',nm,'|Mbwd|',nb)
