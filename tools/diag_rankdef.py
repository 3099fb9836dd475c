"""Diagnostics for the rank-deficient solve (csrc/lsq.cu): which path ran, the null-direction coefficient of the
reference's 1000 x 1001 KAT, phase-mode dependence, and timing of cfg1."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lpvspectral_jl_b200 as lp
from lpvspectral_jl_b200 import _lib as L
from oracle import lpvs_oracle as o

ctx = lp.Context(0)
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
t = np.arange(1000) * 0.1
y = np.sin(2 * np.pi * t)
xs, _ = o.ls_spectral(y, t, mode="literal")
xq, _ = o.ls_spectral(y, t, mode="qr")
for mode in (0, 2):
    ctx.set_option(L.OPT_PHASE_MODE, mode)
    x, f, info = lp.ls_spectral(y, t, ctx=ctx, return_info=True)
    print(f"KAT phase={mode} info={info} x[500]={x[500]} gpu-qr {rel(x, xq):.2e} gpu-svd {rel(x, xs):.2e} "
          f"without x[500]: {rel(x[:500], xq[:500]):.2e}; svd x[500]={xs[500]} qr x[500]={xq[500]}")
ctx.set_option(L.OPT_PHASE_MODE, 0)
rng = np.random.default_rng(1)
N = 4096
t = np.sort(10 * rng.random(N))
y = np.sin(2 * np.pi * 20 * t) + 0.5 * np.cos(2 * np.pi * 55 * t + 1) + 0.1 * rng.standard_normal(N)
for nf in (1024, 2048):
    f = lp.default_freqs(t)[:nf]
    for rep in range(3):
        t0 = time.perf_counter()
        x, _, info = lp.ls_spectral(y, t, f, ctx=ctx, return_info=True)
        wall = time.perf_counter() - t0
    print(f"N=4096 Nf={nf}: info={info} call {ctx.last_call_ms():.3f} ms wall {wall*1e3:.3f} ms gram {ctx.gram_timing()[0]:.3f} ms")
# dataflow vs grid-barrier TRSV on the same well-conditioned problem
f = lp.default_freqs(t)[:1024]
for flow in (1, 0):
    ctx.set_option(L.OPT_TRSV_FLOW, flow)
    x, _, info = lp.ls_spectral(y, t, f, ctx=ctx, return_info=True)
    if flow:
        x1 = x
    print(f"trsv_flow={flow}: info={info} call {ctx.last_call_ms():.3f} ms" + ("" if flow else f" rel diff {rel(x1, x):.2e}"))
ctx.set_option(L.OPT_TRSV_FLOW, 1)
