"""Short cfg3 ADMM run (setup + one fixed-iteration loop launch) for profiling: python tools/admm_run.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = lp.Context(0)
t, y, f = bench.make_cfg3()
h = C.c_void_p()
ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, y.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), len(y),
                                           f.ctypes.data_as(C.c_void_p), len(f), None, L.PROX_L1, 0.1, 0.05, None, 0,
                                           0.0, C.byref(h)))
s = lp.ADMM(ctx, h)
s.step(iters, 0.0)
ms, bpi = s.timing()
print(f"cfg3: {iters} iterations in {ms:.2f} ms -> {iters / ms * 1e3:.0f} it/s, {bpi * iters / ms / 1e6:.0f} GB/s algorithmic")
s.free()
