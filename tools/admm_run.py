"""Short ADMM run (setup + one fixed-iteration loop launch) for profiling: python tools/admm_run.py [iters] [cfg3|cfg4]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
ctx = lp.Context(0)
h = C.c_void_p()
if cfg == "cfg3":
    t, y, f = bench.make_cfg3()
    ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, y.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                               len(y), f.ctypes.data_as(C.c_void_p), len(f), None, L.PROX_L1, 0.1,
                                               0.05, None, 0, 0.0, C.byref(h)))
else:
    from oracle import lpvs_oracle as o  # signal generator only

    N = 20000
    Y, V, X = o.generate_lpv_signal(N, seed=4)
    w = 2 * np.pi * np.arange(1, 65) * 0.4
    yv, Xv, Vv, wv = map(lp._api._f64, (Y, X, V, w))
    ctx.check(ctx.lib.lpvs_admm_create_lpv(ctx.h, lp._api._ptr(yv), lp._api._ptr(Xv), lp._api._ptr(Vv), N,
                                           lp._api._ptr(wv), 64, 50, 0, 1, 0.1, 0.05, C.byref(h)))
s = lp.ADMM(ctx, h)
s.step(iters, 0.0)
ms, bpi = s.timing()
print(f"{cfg}: {iters} iterations in {ms:.2f} ms -> {iters / ms * 1e3:.0f} it/s, "
      f"{bpi * iters / ms / 1e6:.0f} GB/s algorithmic")
s.free()
