"""BASELINE.json full sizes through size-independent properties (-m gpu): linearity in y, coherence(y,y) == 1,
window-range additivity, two independent x-update kernels agreeing, and the largest case the oracle still
finishes in seconds (n = 4095 unknowns)."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_cfg2_full_size_properties(ctx):
    import bench
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    t, y, f, n = bench.make_cfg2()
    S1, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    S3, _ = lp.ls_windowpsd(3.0 * y, t, f, nw=1024, window_func=lp.hanning, ctx=ctx)
    assert rel(S3, 9.0 * S1) <= 1e-13  # |x|^2 is quadratic in y
    assert set(np.argsort(-S1)[:2].tolist()) == {40, 100}  # the two planted tones
    assert np.all(S1 > 0) and np.all(np.isfinite(S1))
    W = lp.hanning(n)
    K = lp.window_count(len(y), n, -1)
    assert K == 2047
    a = lp.window_sums(L.WIN_PSD, y, None, t, f, W, n, n >> 1, 1e-10, 0, 1000, ctx=ctx)
    b = lp.window_sums(L.WIN_PSD, y, None, t, f, W, n, n >> 1, 1e-10, 1000, K, ctx=ctx)
    assert rel(lp.window_finalize(L.WIN_PSD, a + b, len(f), K), S1) <= 1e-13
    # last window vs the oracle (phase arithmetic is hardest there: largest absolute time)
    sl = slice((K - 1) * (n >> 1), (K - 1) * (n >> 1) + n)
    xg, _ = lp.ls_spectral(y[sl], t[sl], f, W, ctx=ctx)
    xr, _ = o.ls_spectral(y[sl], t[sl], f, W, mode="literal")
    assert rel(xg, xr) <= 1e-9


def test_cfg5_coherence_properties(ctx):
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(5)
    NS, n = 1 << 22, 4096
    t = np.sort(10 * rng.random(NS))
    fs = 1.0 / np.mean(np.diff(t))
    f = np.arange(512) * 2 * fs / n
    y = np.sin(2 * np.pi * f[40] * t) + 0.5 * np.cos(2 * np.pi * f[100] * t + 1) + 0.1 * rng.standard_normal(NS)
    u = 0.7 * np.roll(y, 5) + 0.5 * rng.standard_normal(NS)
    C1, _ = lp.ls_cohere(y, y, t, f, nw=NS // n, ctx=ctx)
    assert np.all(C1 == 1)  # test/runtests.jl:207-208 at 2047 windows x 1023 unknowns
    C, _ = lp.ls_cohere(y, u, t, f, nw=NS // n, ctx=ctx)
    assert np.all((C >= 0) & (C <= 1 + 1e-12))
    Cs, _ = lp.ls_cohere(u, y, t, f, nw=NS // n, ctx=ctx)  # coherence is symmetric in its arguments
    assert rel(Cs, C) <= 1e-12


def test_cfg3_admm_full_size_variants_agree(ctx):
    """N=16384, Nreg=16383: the SYMV and GEMV x-update kernels are independent implementations."""
    import bench
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    t, y, f = bench.make_cfg3()
    out = {}
    for mode in (1, 0):
        ctx.set_option(L.OPT_ADMM_SYMV, mode)
        try:
            x, _, info = lp.ls_sparse_spectral(y, t, f, lam=0.1, iters=400, tol=0.0, printerval=10 ** 9, ctx=ctx,
                                               return_info=True)
        finally:
            ctx.set_option(L.OPT_ADMM_SYMV, -1)
        out[mode] = info
    assert out[0]["iters"] == out[1]["iters"] == 400
    assert set(np.flatnonzero(out[0]["z"])) == set(np.flatnonzero(out[1]["z"]))
    assert rel(out[1]["z"], out[0]["z"]) <= 1e-10
    assert abs(out[1]["residual"] - out[0]["residual"]) <= 1e-9 * out[0]["residual"]
    z = out[1]["z"]
    tones = [300, 1200, 2500, 4000, 6000]
    assert all(abs(z[k]) + abs(z[len(f) - 1 + k]) > 0 for k in tones)


def test_admm_largest_oracle_size(ctx):
    """n = 4095 unknowns (N=4096): the oracle's exact-x-update mode still finishes in seconds."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(3)
    N = 4096
    t = np.sort(10 * rng.random(N))
    f = lp.default_freqs(t)[: N // 2]
    y = sum(np.sin(2 * np.pi * f[k] * t + i) for i, k in enumerate([30, 120, 250, 400, 600])) \
        + 0.1 * rng.standard_normal(N)
    kw = dict(lam=0.1, iters=150, tol=1e-9)
    x, _, info = lp.ls_sparse_spectral(y, t, f, ctx=ctx, return_info=True, **kw)
    xr, _, ri = o.ls_sparse_spectral(y, t, f, mode="gram", return_info=True, printerval=10 ** 9, **kw)
    assert info["iters"] == ri["iters"]
    assert set(np.flatnonzero(info["z"])) == set(np.flatnonzero(ri["z"]))
    assert rel(info["z"], ri["z"]) <= 1e-9
    og = o.sparse_objective(ri["A"], y, info["z"], o.NormL1(0.1))
    orf = o.sparse_objective(ri["A"], y, ri["z"], o.NormL1(0.1))
    assert abs(og - orf) <= 1e-8 * max(1.0, abs(orf))


def test_cfg4_group_lasso_full_size_variants_agree(ctx):
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    Y, V, X = o.generate_lpv_signal(20000, seed=4)
    w = 2 * np.pi * np.arange(1, 65) * 0.4
    out = {}
    for mode in (1, 0):
        ctx.set_option(L.OPT_ADMM_SYMV, mode)
        try:
            se, info = lp.ls_sparse_spectral_lpv(Y, X, V, w, 50, lam=0.1, iters=300, tol=0.0, printerval=10 ** 9,
                                                 ctx=ctx, return_info=True)
        finally:
            ctx.set_option(L.OPT_ADMM_SYMV, -1)
        out[mode] = (se, info)
    assert rel(out[1][1]["z"], out[0][1]["z"]) <= 1e-10
    p = lp.psd(out[1][0])
    assert set((np.argsort(-p)[:3] + 1).tolist()) == {5, 25, 50}  # 2, 10, 20 Hz on the 0.4 Hz grid


def test_cfg2_windowed_sparse_full_size_properties(ctx):
    """cfg2's record through the batched windowed-sparse path (2047 device ADMM problems in one pass):
    every window stops on its own test, only the planted tones survive the L1 threshold, and single windows of the
    batched kernel (one CTA per window) reproduce the single-problem solver (cooperative multi-CTA kernel) -- two
    independent x-update kernels -- including the iteration count, at the first, a middle and the last window."""
    import bench
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    t, y, f, n = bench.make_cfg2()
    kw = dict(iters=3000, tol=1e-9)
    S1, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, estimator=lp.ls_sparse_spectral,
                            proxg=lp.NormL1(2.0), ctx=ctx, **kw)
    its = ctx.last_window_iters
    assert its.shape == (2047, 1) and 2 <= its.min() and its.max() < 3000
    assert set(np.flatnonzero(S1).tolist()) == {40, 100}  # only the two planted tones survive the threshold
    W = lp.hanning(n)
    hop = n >> 1
    for k in (0, 1023, 2046):
        sl = slice(k * hop, k * hop + n)
        x, _, info = lp.ls_sparse_spectral(y[sl], t[sl], f, W, proxg=lp.NormL1(2.0), ctx=ctx, return_info=True,
                                           printerval=10 ** 9, **kw)
        s, i, _ = lp.window_sparse_sums(L.WIN_PSD, y, None, t, f, W, n, hop, lp.NormL1(2.0), 0.05, 3000, 1e-9, k, k + 1,
                                        ctx=ctx, return_info=True)
        assert int(i[0, 0]) == info["iters"] == int(its[k, 0])
        assert rel(s, x.real ** 2 + x.imag ** 2) <= 1e-10
