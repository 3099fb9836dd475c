"""CPU study (numpy long double): how far the exact-phase / ideal-grid solution of cfg5a windows is from the reference-rounded one,
and whether removing the separable part of the phase difference (dw_k t_s) helps (it does not: DESIGN.md section 3a)."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lpvs_oracle as o
rng = np.random.default_rng(5)
NS, n = 1 << 24, 4096
t = np.sort(10 * rng.random(NS))
fs = 1.0 / np.mean(np.diff(t))
f = np.arange(512) * 2 * fs / n
y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
W = o.hanning(n)
LD=np.longdouble
twopi=2*LD(np.pi) if False else LD('6.283185307179586476925286766559005768')
def sol(A, yy):
    AtW=A.T*W
    G=AtW@A + 1e-10*np.eye(A.shape[1])
    return np.linalg.solve(G, AtW@yy), np.linalg.cond(G)
for k in (4166, 614, 8000, 100):
    sl=slice(k*2048, k*2048+n); tt=t[sl]; yy=y[sl]
    w=2*np.pi*f                     # fl(2 pi f)
    ph_ref=np.outer(tt, w)          # fl(w t)
    Aref=np.hstack([np.cos(ph_ref), -np.sin(ph_ref[:,1:])])/np.sqrt(2*512)
    df=(f[-1]-f[0])/(len(f)-1)
    fid=LD(f[0])+LD(df)*np.arange(512).astype(LD)
    def build(ph_ld):
        turns=ph_ld/twopi; r=(turns-np.rint(turns)).astype(np.float64)
        return np.hstack([np.cos(2*np.pi*r), -np.sin(2*np.pi*r[:,1:])])/np.sqrt(2*512)
    ph_ideal=np.outer(tt.astype(LD), twopi*fid)
    ph_sep=np.outer(tt.astype(LD), w.astype(LD))   # exact product with the rounded angular frequency
    ph_actual=np.outer(tt.astype(LD), twopi*f.astype(LD))  # exact phases of the given doubles f
    xr,c=sol(Aref,yy); xi,_=sol(build(ph_ideal),yy); xs,_=sol(build(ph_sep),yy); xa,_=sol(build(ph_actual),yy)
    rel=lambda a,b: np.linalg.norm(a-b)/np.linalg.norm(b)
    print(k,'cond %.1e'%c,'ideal-grid %.1e  exact-f %.1e  sep(+dw t) %.1e'%(rel(xi,xr),rel(xa,xr),rel(xs,xr)))
