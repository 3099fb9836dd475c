"""GPU parity: Fourier LS estimators through the C ABI vs the oracle (-m gpu).

Tolerances: coefficients <= 1e-9 relative l2 on well-conditioned FP64 problems (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def signal(N, seed, tones=((20.0, 1.0, 0.0), (55.0, 0.5, 1.0)), noise=0.1, T=10.0):
    rng = np.random.default_rng(seed)
    t = np.sort(T * rng.random(N))
    y = sum(a * np.cos(2 * np.pi * f * t + p) for f, a, p in tones) + noise * rng.standard_normal(N)
    return t, y


@pytest.mark.parametrize("phase", [1, 2, 3, 4])
@pytest.mark.parametrize("N,Nf,zero", [(1000, 100, True), (777, 70, False), (300, 129, True)])
def test_gram_matches_oracle(ctx, phase, N, Nf, zero):
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    t, y = signal(N, 11)
    f = (np.arange(Nf) + (0 if zero else 1)) * 0.37
    W = 0.5 + np.random.default_rng(5).random(N)
    ctx.set_option(L.OPT_PHASE_MODE, phase)
    try:
        G, b = lp.gram_fourier(t, f, W, y, ctx=ctx)
    finally:
        ctx.set_option(L.OPT_PHASE_MODE, 0)
    A, _ = o.get_fourier_regressor(t, f)
    Gr = (A.T * W) @ A
    br = (A.T * W) @ y
    assert np.abs(G - Gr).max() <= 2e-13 * np.abs(Gr).max()
    assert np.abs(b - br).max() <= 2e-13 * np.abs(br).max()
    assert np.array_equal(G, G.T)


def test_chain_ref_reproduces_reference_phase_rounding(ctx):
    """SURVEY H3 at its hardest: large absolute time and high frequency (phi = 2 pi f t ~ 2.6e7 rad, the scale of cfg5a's
    last windows), where the reference's fl(fl(2 pi f) t) is off the true phase by up to 2.9e-9 rad.  The Gram matrix of the
    default mode (chain_ref) must match the oracle's (reference rounding) as closely as the per-element mode does; the
    exact-phase chain must differ by about the phase rounding, and no more."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    rng = np.random.default_rng(12)
    N, Nf = 4096, 200
    t = 9.99 + np.sort(2.4e-3 * rng.random(N))
    fs = N / 2.4e-3
    f = np.arange(Nf) * (fs / N) * 5.0
    W = o.hanning(N)
    y = rng.standard_normal(N)
    A, _ = o.get_fourier_regressor(t, f)
    Gr, br = (A.T * W) @ A, (A.T * W) @ y
    err = {}
    for mode in (L.PHASE_AUTO, L.PHASE_DIRECT, L.PHASE_CHAIN):
        ctx.set_option(L.OPT_PHASE_MODE, mode)
        try:
            G, b = lp.gram_fourier(t, f, W, y, ctx=ctx)
        finally:
            ctx.set_option(L.OPT_PHASE_MODE, 0)
        err[mode] = max(np.abs(G - Gr).max() / np.abs(Gr).max(), np.abs(b - br).max() / np.abs(br).max())
    phimax = 2 * np.pi * f[-1] * t[-1]
    print(f"phi_max {phimax:.2e} rad, phase rounding {phimax * 2.2e-16 / 2:.1e}; Gram error vs oracle: auto {err[0]:.1e} "
          f"direct {err[2]:.1e} exact-phase chain {err[1]:.1e}")
    assert err[L.PHASE_AUTO] <= 5e-13 and err[L.PHASE_DIRECT] <= 5e-13
    assert 1e-11 < err[L.PHASE_CHAIN] <= 4 * phimax * 2.2e-16


def test_parity1_unweighted_and_weighted(ctx):
    """SURVEY 8(d) PARITY-1: N=4096, f=default_freqs(t)[:1024], cond(A)=285."""
    import lpvspectral_jl_b200 as lp

    t, y = signal(4096, 1)
    f = o.default_freqs(t)[:1024]
    x, _ = lp.ls_spectral(y, t, f, ctx=ctx)
    xr, _ = o.ls_spectral(y, t, f, mode="literal")
    assert rel(x, xr) <= TOL
    # Hann weights halve the resolution: spacing 2/T keeps the weighted problem well conditioned
    # (cond(A'WA) = 1e3; at spacing 1/T it is 4e8 and even the oracle's two CPU modes differ by 2e-9).
    W = o.hanning(4096)
    f2 = f[::2]
    xw, _ = lp.ls_spectral(y, t, f2, W, ctx=ctx)
    xwr, _ = o.ls_spectral(y, t, f2, W, mode="literal")
    assert rel(xw, xwr) <= TOL
    # ill-conditioned variant: bounded by cond * phase-rounding differences, documented looser bar
    xw, _ = lp.ls_spectral(y, t, f, W, ctx=ctx)
    xwr, _ = o.ls_spectral(y, t, f, W, mode="literal")
    assert rel(xw, xwr) <= 1e-7


def test_reference_kats(ctx):
    """test/runtests.jl:186-208 through the GPU path."""
    import lpvspectral_jl_b200 as lp

    t = np.arange(1000) * 0.1
    f = lp.default_freqs(t)
    y = np.sin(2 * np.pi * t)
    x, fr, info = lp.ls_spectral(y, t, ctx=ctx, return_info=True)  # rank-deficient 1000x1001 (H1)
    a = x.real ** 2 + x.imag ** 2
    assert a.argmax() + 1 == 101 and abs(a.max() - 2.0 * len(fr)) < 1e-4
    x, _ = lp.ls_spectral(y, t, f, np.ones(len(y)), ctx=ctx)
    a = x.real ** 2 + x.imag ** 2
    assert a.argmax() + 1 == 101 and abs(a.max() - 2.0 * len(f)) < 1e-4
    S, fr = lp.ls_windowpsd(y, t, noverlap=0, ctx=ctx)
    assert S.argmax() + 1 == 13
    S, fr = lp.ls_windowpsd(y, t, nw=16, noverlap=0, ctx=ctx)
    assert np.abs(S).argmax() + 1 == 7
    S, fr = lp.ls_windowcsd(y, y, t, noverlap=0, ctx=ctx)
    assert np.abs(S).argmax() + 1 == 11 and abs(np.abs(S).max() - 2.0 * len(fr)) < 1e-4
    Cxy, _ = lp.ls_cohere(y, y, t, ctx=ctx)
    assert np.all(Cxy == 1)


@pytest.mark.parametrize("kind", ["psd", "csd", "cohere"])
def test_windowed_parity(ctx, kind):
    import lpvspectral_jl_b200 as lp

    N, nw = 6000, 10
    t, y = signal(N, 3, tones=((30.0, 1.0, 0.0), (70.0, 0.7, 0.4)))
    rng = np.random.default_rng(9)
    u = 0.7 * np.roll(y, 3) + 0.5 * rng.standard_normal(N)
    n = N // nw
    fs = 1.0 / np.mean(np.diff(t))
    f = np.arange(40) * 2 * fs / n
    if kind == "psd":
        S, _ = lp.ls_windowpsd(y, t, f, nw=nw, window_func=lp.hanning, ctx=ctx)
        Sr, _ = o.ls_windowpsd(y, t, f, nw=nw, window_func=o.hanning)
    elif kind == "csd":
        S, _ = lp.ls_windowcsd(y, u, t, f, nw=nw, window_func=lp.hanning, ctx=ctx)
        Sr, _ = o.ls_windowcsd(y, u, t, f, nw=nw, window_func=o.hanning)
    else:
        S, _ = lp.ls_cohere(y, u, t, f, nw=nw, ctx=ctx)
        Sr, _ = o.ls_cohere(y, u, t, f, nw=nw)
    assert rel(S, Sr) <= TOL


def test_errors(ctx):
    import lpvspectral_jl_b200 as lp

    t, y = signal(100, 2)
    with pytest.raises(ValueError):
        lp.ls_spectral(y, t, np.array([1.0, 0.0, 2.0]), ctx=ctx)
    with pytest.raises(ValueError):
        lp.ls_spectral(y[:-1], t, np.array([1.0, 2.0]), ctx=ctx)


def test_window_ranges_add_up(ctx):
    """Sharding unit: sums over disjoint window ranges add to the full-range sums (what ranks all-reduce)."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    N, nw = 8192, 16
    t, y = signal(N, 21)
    u = np.roll(y, 1)
    n = N // nw
    f = np.arange(24) * 2.0 / (t[n] - t[0])
    W = lp.hanning(n)
    K = lp.window_count(N, n, -1)
    for kind in (L.WIN_PSD, L.WIN_CSD, L.WIN_COHERE):
        uu = None if kind == L.WIN_PSD else u
        full = lp.window_sums(kind, y, uu, t, f, W, n, n >> 1, 1e-10, 0, K, ctx=ctx)
        a = lp.window_sums(kind, y, uu, t, f, W, n, n >> 1, 1e-10, 0, 11, ctx=ctx)
        b = lp.window_sums(kind, y, uu, t, f, W, n, n >> 1, 1e-10, 11, K, ctx=ctx)
        assert np.allclose(a + b, full, rtol=1e-12, atol=1e-300)
    # small batches (forces several factorisation batches) give bit-identical sums
    ctx.set_option(L.OPT_WINDOW_BATCH, 5)
    try:
        small = lp.window_sums(L.WIN_PSD, y, None, t, f, W, n, n >> 1, 1e-10, 0, K, ctx=ctx)
    finally:
        ctx.set_option(L.OPT_WINDOW_BATCH, 0)
    assert np.array_equal(small, lp.window_sums(L.WIN_PSD, y, None, t, f, W, n, n >> 1, 1e-10, 0, K, ctx=ctx))


def test_rowsharded_single_rank_equals_ls_spectral(ctx):
    """lpvs_gram_partial_dev + lpvs_solve_packed_dev (the row-sharded path) at world size 1, incl. sample splits."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _dist as D

    N = 100000
    t, y = signal(N, 22)
    u = np.roll(y, 3)
    f = np.arange(96) * 0.5
    W = 0.5 + np.random.default_rng(1).random(N)
    xs = D.ls_spectral_rowsharded(y, t, f, W, u=u, lam=1e-10, ctx=ctx)
    xy, _ = lp.ls_spectral(y, t, f, W, ctx=ctx)
    xu, _ = lp.ls_spectral(u, t, f, W, ctx=ctx)
    assert rel(xs[0], xy) <= 1e-12 and rel(xs[1], xu) <= 1e-12
    A, _ = o.get_fourier_regressor(t[:20000], f)
    xr, _ = o.ls_spectral(y[:20000], t[:20000], f, W[:20000], mode="gram")
    xg = D.ls_spectral_rowsharded(y[:20000], t[:20000], f, W[:20000], lam=1e-10, ctx=ctx)
    assert rel(xg, xr) <= TOL


def test_ragged_and_tiny_inputs(ctx):
    """Edge cases: n not a multiple of the 32-sample chunk, Nf = 1, a window longer than the signal (K = 0)."""
    import lpvspectral_jl_b200 as lp

    t, y = signal(333, 23)
    f = np.array([0.0])
    x, _ = lp.ls_spectral(y, t, f, ctx=ctx)
    xr, _ = o.ls_spectral(y, t, f, mode="literal")
    assert rel(x, xr) <= TOL
    f = np.arange(1, 8) * 0.7
    x, _ = lp.ls_spectral(y, t, f, np.ones(333), ctx=ctx)
    xr, _ = o.ls_spectral(y, t, f, np.ones(333), mode="literal")
    assert rel(x, xr) <= TOL
    S, _ = lp.ls_windowpsd(y, t, f, nw=3, noverlap=7, ctx=ctx)  # n = 111 (ragged chunks), K = 3
    Sr, _ = o.ls_windowpsd(y, t, f, nw=3, noverlap=7)
    assert rel(S, Sr) <= TOL
    assert lp.window_count(10, 20, 0) == 0


def test_nonfinite_inputs_are_reported(ctx):
    """LPVS_E_NONFINITE: NaN/Inf in a host input array is detected on the device at upload time."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    t, y = signal(512, 30)
    f = np.arange(1, 20) * 0.5
    yb = y.copy()
    yb[100] = np.nan
    with pytest.raises(lp.LpvsError) as ei:
        lp.ls_spectral(yb, t, f, ctx=ctx)
    assert ei.value.code == L.E_NONFINITE
    tb = t.copy()
    tb[7] = np.inf
    with pytest.raises(lp.LpvsError) as ei:
        lp.ls_windowpsd(y, tb, f, nw=4, ctx=ctx)
    assert ei.value.code == L.E_NONFINITE
    x, _ = lp.ls_spectral(y, t, f, ctx=ctx)  # the context stays usable
    assert np.all(np.isfinite(x))


def test_float32_signatures(ctx):
    """the reference's signatures are eltype-generic (e.g. src/lasso.jl:85).  A Float32 signal gives
    Float32 / ComplexF32 results; the arithmetic is the FP64 device path on the up-converted inputs, so the result must
    equal the FP64 answer for the same (single-precision) inputs rounded to single."""
    import lpvspectral_jl_b200 as lp

    t, y = signal(2048, 9)
    t32, y32 = t.astype(np.float32), y.astype(np.float32)
    f = np.arange(0, 60) * 0.9
    x32, _ = lp.ls_spectral(y32, t32, f, ctx=ctx)
    x64, _ = lp.ls_spectral(y32.astype(np.float64), t32.astype(np.float64), f, ctx=ctx)
    assert x32.dtype == np.complex64 and x64.dtype == np.complex128
    assert np.array_equal(x32, x64.astype(np.complex64))
    S32, _ = lp.ls_windowpsd(y32, t32, f, nw=4, window_func=lp.hanning, ctx=ctx)
    S64, _ = lp.ls_windowpsd(y32.astype(np.float64), t32.astype(np.float64), f, nw=4, window_func=lp.hanning, ctx=ctx)
    assert S32.dtype == np.float32 and np.array_equal(S32, S64.astype(np.float32))
    z32, _ = lp.ls_sparse_spectral(y32, t32, f[1:], lam=0.2, iters=300, tol=1e-9, ctx=ctx)
    assert z32.dtype == np.complex64
    # and against the oracle run in FP64 on the same single-precision inputs: single-precision agreement
    xr, _ = o.ls_spectral(y32.astype(np.float64), t32.astype(np.float64), f, mode="literal")
    assert rel(x32, xr) <= 1e-6
