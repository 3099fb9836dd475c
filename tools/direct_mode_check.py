"""GRAM_DIRECT (non-uniform frequency grid, per-element sincos with the reference's phase rounding) vs GRAM_CHAIN
on a cfg2-shaped problem: speed of the Gram kernel and agreement of the two synthesis paths."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

ctx = lp.Context(0)
t, y, f, n = bench.make_cfg2(nsamp=1 << 21, nw=512)
out = {}
for mode, name in ((L.PHASE_CHAIN, "chain"), (L.PHASE_DIRECT, "direct")):
    ctx.set_option(L.OPT_PHASE_MODE, mode)
    for rep in range(2):
        S, _ = lp.ls_windowpsd(y, t, f, nw=512, window_func=lp.hanning, ctx=ctx)
    ms, nl, fl = ctx.gram_timing()
    out[name] = S
    print(f"{name}: gram {ms:.2f} ms, {fl / ms / 1e9:.2f} TFLOP/s useful ({fl / ms / 1e9 / 36.962 * 100:.1f} % of DMMA peak)")
ctx.set_option(L.OPT_PHASE_MODE, 0)
print("chain vs direct PSD rel diff", np.linalg.norm(out["chain"] - out["direct"]) / np.linalg.norm(out["direct"]))
