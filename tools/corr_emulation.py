import numpy as np, sys, os
sys.path.insert(0, "/root/repo")
from oracle import lpvs_oracle as o
rng = np.random.default_rng(5)
NS, n = 1 << 24, 4096
t = np.sort(10 * rng.random(NS))
fs = 1.0 / np.mean(np.diff(t))
Nf=512
f = np.arange(Nf) * 2 * fs / n
y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
W = o.hanning(n)
LD=np.longdouble
twopi=LD('6.283185307179586476925286766559005768')
dd=1/np.sqrt(2*Nf)
lam=1e-10
def solve(G,b):
    return np.linalg.solve(G+lam*np.eye(G.shape[0]), b)
rel=lambda a,b: np.linalg.norm(a-b)/np.linalg.norm(b)
df=(f[-1]-f[0])/(Nf-1)
fid=LD(f[0])+LD(df)*np.arange(Nf).astype(LD)
w=2*np.pi*f
dw=(w.astype(LD)-twopi*fid).astype(np.float64)
for k in (4166, 614, 100):
    sl=slice(k*2048, k*2048+n); tt=t[sl]; yy=y[sl]
    p=np.outer(tt,w)  # fl(w t)
    cr,sr=np.cos(p),-np.sin(p)
    Aref=np.hstack([cr,sr[:,1:]])*dd
    Gref=(Aref.T*W)@Aref; bref=(Aref.T*W)@yy
    xr=solve(Gref,bref)
    turns=np.outer(tt.astype(LD),fid); r=(turns-np.rint(turns)).astype(np.float64)
    ci,si=np.cos(2*np.pi*r),-np.sin(2*np.pi*r)
    Aid=np.hstack([ci,si[:,1:]])*dd
    Gid=(Aid.T*W)@Aid; bid=(Aid.T*W)@yy
    xi=solve(Gid,bid)
    # eps = phi_ref - theta_ideal = (p - w t) + dw t  (exact product error via long double)
    e=(np.outer(tt.astype(LD),w.astype(LD))-p.astype(LD)).astype(np.float64)  # w t - p
    eps=dw[None,:]*tt[:,None]-e
    print(k,'max|eps| %.2e'%np.abs(eps).max())
    S=2.0**np.floor(np.log2(32000/np.abs(eps).max()))  # fixed-point scale
    epsq=np.rint(eps*S)   # integer fixed point
    for name,dt in (("fp16",np.float16),("f32",np.float32)):
        # elements: D_c = W eps ms ; D_s = -W eps c ; B = (c, ms); use cos/sin of p (reference phase) in fp32
        c32=np.cos(p).astype(np.float32); ms32=(-np.sin(p)).astype(np.float32)
        de=(epsq.astype(np.float32)*W[:,None].astype(np.float32))
        Dc=(de*ms32).astype(dt).astype(np.float32); Ds=(-de*c32).astype(dt).astype(np.float32)
        Bc=c32.astype(dt).astype(np.float32); Bs=ms32.astype(dt).astype(np.float32)
        D=np.hstack([Dc,Ds[:,1:]]); B=np.hstack([Bc,Bs[:,1:]])
        M=D.T@B
        dG=(M+M.T).astype(np.float64)/S*dd*dd
        db=(np.hstack([de*ms32,(-de*c32)[:,1:]]).astype(np.float64).T@yy)/S*dd
        xc=solve(Gid+dG,bid+db)
        xc2=solve(Gid+dG,bid)
        print('  %s: uncorrected %.2e  corrected %.2e (G only: %.2e)  |dG-(Gref-Gid)|/|Gref-Gid| %.2e  |Gref-Gid|/|G| %.2e'%(name,rel(xi,xr),rel(xc,xr),rel(xc2,xr),
              np.linalg.norm(dG-(Gref-Gid))/np.linalg.norm(Gref-Gid), np.linalg.norm(Gref-Gid)/np.linalg.norm(Gref)))
