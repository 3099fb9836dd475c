import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import lpvspectral_jl_b200 as lp
seed = int(sys.argv[1])
rng = np.random.default_rng(4000 + seed); N = int(rng.integers(40, 900)); ratio = [0.55, 0.8, 1.0, 1.3, 2.0][seed % 5]
Nf = max(2, int(ratio * N / 2)); zero = bool(seed % 2)
t = np.sort(10 * rng.random(N)); fs = N / 10.0
f = (np.arange(Nf) + (0 if zero else 1)) * (fs / 2 / Nf) * (0.9 if seed % 3 else 1.0)
y = np.sin(2 * np.pi * f[Nf // 3] * t) + 0.3 * rng.standard_normal(N); lam = [1e-10, 1e-8, 1e-6][seed % 3]
ctx = lp.Context(0)
try:
    x, _, info = lp.ls_spectral(y, t, f, lam=lam, ctx=ctx, return_info=True)
    print("ok info", info)
except Exception as e:
    print("ERR", e)
