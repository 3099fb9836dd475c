"""Determinism check: the windowed estimators must give bit-identical results run to run and for any device batch size
(python tools/det_check.py on the GPU box)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import lpvspectral_jl_b200 as lp
from lpvspectral_jl_b200 import _lib as L
ctx = lp.Context(0)
rng = np.random.default_rng(5)
NS = 1 << 20
t = np.sort(10 * rng.random(NS)); n = 4096
fs = 1.0 / np.mean(np.diff(t)); f = np.arange(512) * 2 * fs / n
y = np.sin(2*np.pi*f[40]*t) + 0.5*np.cos(2*np.pi*f[100]*t+1) + 0.1*rng.standard_normal(NS)
u = 0.7*np.roll(y, 5) + 0.5*rng.standard_normal(NS)
res = {}
for b in (0, 768, 2047, 148, 739):
    ctx.set_option(L.OPT_WINDOW_BATCH, b)
    C, _ = lp.ls_cohere(y, u, t, f, nw=NS // n, ctx=ctx)
    S, _ = lp.ls_windowpsd(y, t, f, nw=NS // n, window_func=lp.hanning, ctx=ctx)
    key = (b, len(res))
    res[key] = (C.copy(), S.copy())
    print(b, repr(C[40]), repr(C[100]), repr(S[40]))
k0 = list(res)[0]
for k in res:
    print(k, np.max(np.abs(res[k][0] - res[k0][0])), np.max(np.abs(res[k][1] - res[k0][1]) / np.max(res[k0][1])))
