"""cfg2 windowed PSD in LPVS_PHASE_STRUCTURED, or LPVS_PHASE_STRUCTURED_REF with LPVS_PROFILE_MODE=5 (run under ncu for the
launch list): python tools/structured_profile.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

ctx = lp.Context(0)
ctx.set_option(L.OPT_PHASE_MODE, int(os.environ.get("LPVS_PROFILE_MODE", L.PHASE_STRUCTURED)))
t, y, f, n = bench.make_cfg2()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    S, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    print("call ms", ctx.last_call_ms(), "gram stage ms", ctx.gram_timing()[0])
