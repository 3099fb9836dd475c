"""BASELINE cfg5a (ls_cohere, 2^24 samples per channel, 8191 windows, 512 frequencies) in every phase mode: how far the
whole-record coherence of each mode is from the default mode's.  The yardstick for LPVS_PHASE_STRUCTURED_REF is the spread
between the two modes that carry the reference's phase rounding exactly (default chain_ref vs per-element direct): a few
ill-conditioned windows (cond(A'WA) up to 4.5e7 and beyond) dominate the cross-window sums, and each mode carries its own
cond x eps there.   python tools/cfg5a_mode_spread.py > gpurun_out/cfg5a_mode_spread.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

ctx = lp.Context(0)
t, y, u, f, n = bench.make_cfg5()
NS, Nf = len(t), len(f)
W = lp.hanning(n)
hop = n >> 1
K = lp.window_count(NS, n, hop)
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
info = C.c_int(0)
res = {}
coh = {}
MODES = (("default_chain_ref", L.PHASE_AUTO), ("direct", L.PHASE_DIRECT), ("structured_ref", L.PHASE_STRUCTURED_REF),
         ("chain_exact_phase", L.PHASE_CHAIN), ("structured", L.PHASE_STRUCTURED))
want = set(sys.argv[1:])  # optional subset of mode names (the default mode always runs: it is the yardstick)
for name, mode in MODES:
    if want and name not in want and name != "default_chain_ref":
        continue
    ctx.set_option(L.OPT_PHASE_MODE, mode)
    sums = np.zeros(4 * Nf)
    w0 = time.perf_counter()
    ctx.check(ctx.lib.lpvs_ls_window_sums(ctx.h, L.WIN_COHERE, vp(y), vp(u), vp(t), NS, vp(f), Nf, vp(W), n, hop, bench.LAMBDA, 0,
                                          K, vp(sums), C.byref(info)))
    dt = time.perf_counter() - w0
    coh[name] = lp.window_finalize(L.WIN_COHERE, sums, Nf, K)
    res[name] = {"s_per_pass_cold": dt, "info": info.value}
ctx.set_option(L.OPT_PHASE_MODE, L.PHASE_AUTO)
ref = coh["default_chain_ref"]
for name in coh:
    res[name]["coherence_rel_l2_vs_default"] = float(np.linalg.norm(coh[name] - ref) / np.linalg.norm(ref))
    res[name]["coherence_max_abs_vs_default"] = float(np.abs(coh[name] - ref).max())
print(json.dumps(res))
