"""Phase timeline of the persistent ADMM kernel (diagnostics): python tools/admm_trace.py [cfg3|cfg4] [iters]

Runs the loop with LPVS_ADMM_TRACE set; the kernel stamps clock64 at 8 points of 4 mid-run iterations in every CTA.
Prints, per phase, min / median / max over CTAs in microseconds (SM clock taken as 1.965 GHz)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
TRACE = os.path.join(OUT, "admm_trace.bin")
os.environ["LPVS_ADMM_TRACE"] = TRACE

import ctypes as C  # noqa: E402

import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
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
s.step(50, 0.0)
s.step(iters, 0.0)
ms, bpi = s.timing()
print(f"{cfg}: {iters} iterations in {ms:.2f} ms -> {iters / ms * 1e3:.0f} it/s = {ms / iters * 1e3:.2f} us/it, "
      f"{bpi * iters / ms / 1e6:.0f} GB/s algorithmic")
s.free()
raw = open(TRACE, "rb").read()
ni, grid = np.frombuffer(raw[:8], dtype=np.int32)
tr = np.frombuffer(raw[8:], dtype=np.int64).reshape(ni, grid, 8).astype(np.float64) / 1965.0  # us
names = ["zero+phase1", "publish", "barrier1", "phase2(+group prox)", "barrier2(top-r)", "top-r select", "end_iter(barrier)"]
for it in range(ni):
    d = np.diff(tr[it], axis=1)
    print(f"iteration {it}: total {np.median(tr[it, :, 7] - tr[it, :, 0]):.2f} us (median CTA)")
    for k, nm in enumerate(names):
        col = d[:, k]
        print(f"   {nm:22s} min {col.min():7.2f}  med {np.median(col):7.2f}  max {col.max():7.2f}")
if ni > 1:
    print("start-to-start:", np.median(tr[1:, :, 0] - tr[:-1, :, 0], axis=1))
