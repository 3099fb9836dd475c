"""Full-size runs of the BASELINE.json configs with timing and size-independent property checks.

  python tools/measure_configs.py cfg1 cfg2 cfg3 cfg4 cfg5a [--iters N]

Writes gpurun_out/measure_<cfg>.json.  Properties checked at full size (the oracle cannot run these sizes in
seconds): linearity of the estimator in y, coherence(y,y) == 1, normal-equation residual, ADMM fixed-point/KKT."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _lib as L  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
FP64_PEAK = 36.962
HBM_PEAK = 6554.6


def dump(name, d):
    with open(os.path.join(OUT, f"measure_{name}.json"), "w") as fh:
        json.dump(d, fh, indent=1)
    print(name, json.dumps(d))


def cfg1(ctx, args):
    rng = np.random.default_rng(1)
    N = 4096
    t = np.sort(10 * rng.random(N))
    y = np.sin(2 * np.pi * 20 * t) + 0.5 * np.cos(2 * np.pi * 55 * t + 1) + 0.1 * rng.standard_normal(N)
    f = lp.default_freqs(t)[:2048]
    res = {}
    for rep in range(4):
        t0 = time.perf_counter()
        x, _, info = lp.ls_spectral(y, t, f, ctx=ctx, return_info=True)
        wall = time.perf_counter() - t0
        gms, gl, gfl = ctx.gram_timing()
        res = dict(wall_ms=wall * 1e3, call_ms=ctx.last_call_ms(), gram_ms=gms, gram_launches=gl,
                   gram_tflops=gfl / gms / 1e9, info=info)
    nreg = 4095
    flop = N * nreg * (nreg + 1) + nreg ** 3 / 3 + 2 * nreg ** 2
    res.update(spectra_per_s=1e3 / res["call_ms"], flop_per_spectrum=flop,
               tflops_e2e=flop / res["call_ms"] / 1e9, frac_gram=res["gram_tflops"] / FP64_PEAK,
               note="cond(A)~1e16: info=2 means the shifted-CholeskyQR path of csrc/lsq.cu ran (the reference's SVD-class "
                    "answer, tests/test_gpu_rankdef.py::test_cfg1_full_size)")
    a = x.real ** 2 + x.imag ** 2
    res["peak_index"] = int(a.argmax())
    res["peak_freq"] = float(f[a.argmax()])
    dump("cfg1", res)


def cfg2(ctx, args):
    sys.path.insert(0, ROOT)
    import bench

    t, y, f, n = bench.make_cfg2()
    S1, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    t0 = time.perf_counter()
    S1, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    wall = time.perf_counter() - t0
    gms, gl, gfl = ctx.gram_timing()
    S2, _ = lp.ls_windowpsd(3.0 * y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    C, _ = lp.ls_cohere(y[: 1 << 20], y[: 1 << 20], t[: 1 << 20], f, nw=256, ctx=ctx)
    dump("cfg2", dict(wall_ms=wall * 1e3, call_ms=ctx.last_call_ms(), gram_ms=gms, gram_tflops=gfl / gms / 1e9,
                      windows_per_s=2047 / wall, linearity_rel=float(np.linalg.norm(S2 - 9 * S1) / np.linalg.norm(9 * S1)),
                      peaks=[int(i) for i in np.argsort(-S1)[:2]], cohere_identical_all_one=bool(np.all(C == 1))))


def cfg3(ctx, args):
    rng = np.random.default_rng(3)
    N = args.get("N", 16384)
    t = np.sort(10 * rng.random(N))
    f = lp.default_freqs(t)[: N // 2]
    tones = f[[300, 1200, 2500, 4000, 6000]] if N >= 16384 else f[[3, 12, 25, 40, 60]]
    y = sum(np.sin(2 * np.pi * ft * t + i) for i, ft in enumerate(tones)) + 0.1 * rng.standard_normal(N)
    t0 = time.perf_counter()
    h = lp._api.C.c_void_p()
    yv, tv, fv = lp._api._f64(y), lp._api._f64(t), lp._api._f64(f)
    ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, lp._api._ptr(yv), lp._api._ptr(tv), N, lp._api._ptr(fv), len(fv),
                                               None, L.PROX_L1, 0.1, 0.05, None, 0, 0.0, lp._api.C.byref(h)))
    setup = time.perf_counter() - t0
    gms, gl, gfl = ctx.gram_timing()
    solver = lp.ADMM(ctx, h)
    iters = args.get("iters", 3000)
    solver.step(50, 0.0)  # warm-up
    t0 = time.perf_counter()
    solver.step(iters, 0.0)  # tol = 0: fixed iteration count
    wall = time.perf_counter() - t0
    ms, bpi = solver.timing()
    its_s = iters / (ms * 1e-3)
    res = dict(N=N, n=2 * len(fv) - 1, setup_s=setup, gram_ms=gms, gram_tflops=gfl / gms / 1e9, iters=iters, loop_ms=ms,
               wall_ms=wall * 1e3, iters_per_s=its_s, bytes_per_iter=bpi, gbs=bpi * its_s / 1e9,
               frac_hbm=bpi * its_s / 1e9 / HBM_PEAK, residual=solver.residual)
    # natural stop run
    t0 = time.perf_counter()
    solver.run(iters=30000 + solver.iters, tol=1e-9, printerval=10 ** 9, verbose=False)
    res.update(natural_stop_iters=solver.iters, natural_stop_converged=solver.converged,
               natural_stop_wall_s=time.perf_counter() - t0, final_residual=solver.residual)
    x, z = solver.get()
    zc = solver.result(len(fv))
    nz = np.flatnonzero(np.abs(zc) > 0)
    res["support_size"] = int(len(nz))
    res["support_contains_tones"] = bool(set(np.searchsorted(f, tones)) <= set(nz.tolist()))
    res["x_minus_z"] = float(np.linalg.norm(x - z))
    solver.free()
    dump("cfg3" if N >= 16384 else f"cfg3_N{N}", res)


def cfg4(ctx, args):
    from oracle import lpvs_oracle as o  # signal generator only

    N = args.get("N", 20000)
    Y, V, X = o.generate_lpv_signal(N, seed=4)
    w = 2 * np.pi * np.arange(1, 65) * 0.4
    yv, Xv, Vv, wv = map(lp._api._f64, (Y, X, V, w))
    h = lp._api.C.c_void_p()
    t0 = time.perf_counter()
    ctx.check(ctx.lib.lpvs_admm_create_lpv(ctx.h, lp._api._ptr(yv), lp._api._ptr(Xv), lp._api._ptr(Vv), N,
                                           lp._api._ptr(wv), 64, 50, 0, 1, 0.1, 0.05, lp._api.C.byref(h)))
    setup = time.perf_counter() - t0
    gms, gl, gfl = ctx.gram_timing()
    solver = lp.ADMM(ctx, h)
    solver.step(50, 0.0)
    iters = args.get("iters", 6000)
    solver.step(iters, 0.0)
    ms, bpi = solver.timing()
    its_s = iters / (ms * 1e-3)
    res = dict(N=N, n=6400, setup_s=setup, gram_ms=gms, gram_tflops=gfl / gms / 1e9, iters=iters, loop_ms=ms,
               iters_per_s=its_s, bytes_per_iter=bpi, gbs=bpi * its_s / 1e9, frac_hbm=bpi * its_s / 1e9 / HBM_PEAK,
               residual=solver.residual)
    params = solver.result(64 * 50)
    rp = np.reshape(params, (64, -1), order="F")
    p = np.abs(rp.sum(axis=1)) ** 2
    res["active_freqs"] = [int(i) + 1 for i in np.flatnonzero(p > 0)]
    res["expected_active"] = [5, 25, 50]  # 2, 10, 20 Hz on the 0.4 Hz grid
    solver.free()
    dump("cfg4", res)


def cfg2s(ctx, args):
    """cfg2's record with estimator = ls_sparse_spectral (L1): K windows x `iters` ADMM iterations, batched."""
    rng = np.random.default_rng(2)
    NS = args.get("N", 1 << 22)
    t = np.sort(10 * rng.random(NS))
    n = 4096
    fs = 1.0 / np.mean(np.diff(t))
    f = np.arange(256) * 2 * fs / n
    y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
    iters = args.get("iters", 300)
    K = lp.window_count(NS, n, -1)
    kw = dict(nw=NS // n, window_func=lp.hanning, estimator=lp.ls_sparse_spectral, lam=0.01, iters=iters, tol=0.0,
              ctx=ctx)
    S, _ = lp.ls_windowpsd(y, t, f, **kw)
    t0 = time.perf_counter()
    S, _ = lp.ls_windowpsd(y, t, f, **kw)
    wall = time.perf_counter() - t0
    Sd, _ = lp.ls_windowpsd(y, t, f, nw=NS // n, window_func=lp.hanning, ctx=ctx)
    t0 = time.perf_counter()
    Sd, _ = lp.ls_windowpsd(y, t, f, nw=NS // n, window_func=lp.hanning, ctx=ctx)
    wall_dense = time.perf_counter() - t0
    Np = 512
    res = dict(samples=NS, windows=K, iters=iters, wall_s=wall, wall_dense_s=wall_dense,
               admm_s=wall - wall_dense, window_iters_per_s=K * iters / max(wall - wall_dense, 1e-9),
               # PSD (one channel) streams the lower 128x128 blocks of each window's inverse: nb(nb+1)/2 * 128 KB
               gbs_lower_M=K * iters * 10 * 131072.0 / max(wall - wall_dense, 1e-9) / 1e9,
               peak_bins=[int(i) for i in np.argsort(S)[-2:]], dense_peak_bins=[int(i) for i in np.argsort(Sd)[-2:]])
    dump("cfg2s", res)


def cfg5a(ctx, args):
    rng = np.random.default_rng(5)
    NS = args.get("N", 1 << 24)
    t = np.sort(10 * rng.random(NS))
    n = 4096
    fs = 1.0 / np.mean(np.diff(t))
    f = np.arange(512) * 2 * fs / n
    y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
    u = 0.7 * np.roll(y, 5) + 0.5 * rng.standard_normal(NS)
    nw = NS // n
    t0 = time.perf_counter()
    C, _ = lp.ls_cohere(y, u, t, f, nw=nw, ctx=ctx)
    wall = time.perf_counter() - t0
    gms, gl, gfl = ctx.gram_timing()
    K = lp.window_count(NS, n, -1)
    t0 = time.perf_counter()
    C, _ = lp.ls_cohere(y, u, t, f, nw=nw, ctx=ctx)
    wall = time.perf_counter() - t0
    gms, gl, gfl = ctx.gram_timing()
    res = dict(samples=NS, windows=K, wall_s=wall, call_ms=ctx.last_call_ms(), windows_per_s=K / wall, gram_ms=gms,
               gram_tflops=gfl / gms / 1e9, frac_fp64=gfl / gms / 1e9 / FP64_PEAK,
               coh_at_tones=[float(C[40]), float(C[100])], coh_median=float(np.median(C)),
               in_unit_interval=bool(np.all((C >= 0) & (C <= 1 + 1e-12))))
    dump("cfg5a", res)


def cfg5b(ctx, args):
    """One weighted LS problem with 2^24 rows, 512 freqs, both channels (row-sharded path at world size 1)."""
    from lpvspectral_jl_b200 import _dist as D

    rng = np.random.default_rng(5)
    NS = args.get("N", 1 << 24)
    t = np.sort(10 * rng.random(NS))
    fs = 1.0 / np.mean(np.diff(t))
    f = np.arange(512) * 2 * fs / 4096
    y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
    u = 0.7 * np.roll(y, 5) + 0.5 * rng.standard_normal(NS)
    W = 0.5 + rng.random(NS)
    for rep in range(2):
        t0 = time.perf_counter()
        x = D.ls_spectral_rowsharded(y, t, f, W, u=u, lam=1e-10, ctx=ctx)
        wall = time.perf_counter() - t0
    a = np.abs(x[0]) ** 2
    nreg = 1023
    res = dict(rows=NS, nreg=nreg, wall_s=wall, flop=float(NS) * nreg * (nreg + 1),
               tflops_wall=float(NS) * nreg * (nreg + 1) / wall / 1e12, peaks=[int(i) for i in np.argsort(-a)[:2]])
    # linearity of the solve in the right-hand side: x(y) and x(u) from one factorisation
    x2 = D.ls_spectral_rowsharded(2.0 * y, t, f, W, u=u, lam=1e-10, ctx=ctx)
    res["linearity_rel"] = float(np.linalg.norm(x2[0] - 2 * x[0]) / np.linalg.norm(2 * x[0]))
    res["u_channel_unchanged"] = float(np.linalg.norm(x2[1] - x[1]) / np.linalg.norm(x[1]))
    dump("cfg5b", res)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    kv = {}
    for a in sys.argv[1:]:
        if a.startswith("--") and "=" in a:
            k, v = a[2:].split("=")
            kv[k] = int(v)
    ctx = lp.Context(0)
    for nme in names:
        t0 = time.time()
        globals()[nme](ctx, kv)
        print(f"[{nme}] done in {time.time() - t0:.1f}s", flush=True)
