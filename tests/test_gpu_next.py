"""SURVEY 8(f) rows n3 and n4 through the GPU path (-m gpu): tls_spectral (src/lsfft.jl:87-99) and the window
re-assembly of mapwindows (Base.merge, src/windows.jl:50-70)."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_tls_spectral_reference_kat(ctx):
    """test/runtests.jl:194-195: y = sin(2 pi t), tls_spectral(y,t) with its default frequencies."""
    import lpvspectral_jl_b200 as lp

    t = np.arange(1000) * 0.1
    y = np.sin(2 * np.pi * t)
    x, f, its = lp.tls_spectral(y, t, ctx=ctx, return_info=True)
    assert len(f) == 500
    a = x.real ** 2 + x.imag ** 2
    assert a.argmax() + 1 == 101 and abs(a.max() - 2.0 * len(f)) < 1e-4
    xr, _ = o.tls_spectral(y, t)
    assert rel(x, xr) <= 1e-9
    assert its <= 5  # sigma_min([A y]) = 0 here: one or two steps of inverse iteration


@pytest.mark.parametrize("N,Nf,noise,seed", [(1000, 100, 0.1, 1), (2048, 300, 0.5, 2), (777, 64, 0.02, 3)])
def test_tls_spectral_against_svd(ctx, N, Nf, noise, seed):
    """Noisy irregular data, well-conditioned regressor: the smallest right singular vector of [A y] by inverse iteration on
    the Gram matrix against LAPACK gesvd (the reference's call)."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(seed)
    t = np.sort(10 * rng.random(N))
    f = np.arange(Nf) * (0.25 * N / 10 / Nf)
    y = np.sin(2 * np.pi * f[Nf // 3] * t) + 0.5 * np.cos(2 * np.pi * f[Nf // 2] * t + 1) + noise * rng.standard_normal(N)
    x, _, its = lp.tls_spectral(y, t, f, ctx=ctx, return_info=True)
    xr, _ = o.tls_spectral(y, t, f)
    print(f"tls N={N} Nf={Nf}: {its} inverse-iteration steps, rel {rel(x, xr):.2e}")
    assert rel(x, xr) <= 1e-9
    # total least squares is NOT ordinary least squares on noisy data: the two must differ measurably
    xl, _ = lp.ls_spectral(y, t, f, ctx=ctx)
    assert rel(x, xl) > 1e-6


@pytest.mark.parametrize("N,n,noverlap", [(100, 10, 0), (100, 10, 5), (1003, 64, 17), (50, 50, -1), (40, 64, 0)])
def test_merge_windows_matches_reference_merge(ctx, N, n, noverlap):
    """Base.merge: bit-exact against the oracle's restatement (same summation order), incl. a dropped tail and K = 0."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(N + n)
    nov = n >> 1 if noverlap < 0 else noverlap
    K = o.arraysplit_count(N, n, nov)
    pieces = rng.standard_normal((K, n))
    got = lp.merge_windows(pieces, N, n, noverlap, ctx=ctx)
    ref = o.merge_windows(list(pieces), N, n, nov) if K else np.zeros(N)
    assert np.array_equal(got, ref)


def test_mapwindows_identity_round_trip(ctx):
    """test/runtests.jl:60-61: mapwindows with the identity gives the signal back where windows cover it."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(0)
    y = rng.standard_normal(100)
    t = np.arange(100.0)
    for nov in (0, 5):
        out = lp.mapwindows(lambda yi, ti: yi, y, t, 10, nov, ctx=ctx)
        cover = (o.arraysplit_count(100, 10, nov) - 1) * (10 - nov) + 10
        assert np.array_equal(out[:cover], y[:cover]) and np.all(out[cover:] == 0)


def test_float32_callers_stream_the_admm_inverse_in_single_precision(ctx):
    """the reference's sparse estimators are eltype-generic (src/lasso.jl:85 `AbstractArray{T}`).  A Float32
    signal makes the ADMM loop keep (G + I/mu)^-1 in single precision (LPVS_OPT_ADMM_M32: half the bytes per iteration,
    double accumulation): results within single-precision distance of the Float64 run, same support, half the bytes."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(3)
    N = 4096
    t = np.sort(10 * rng.random(N))
    f = lp.default_freqs(t)[: N // 2]
    y = sum(np.sin(2 * np.pi * f[k] * t + i) for i, k in enumerate([30, 120, 250, 400, 600])) + 0.1 * rng.standard_normal(N)
    kw = dict(lam=0.1, iters=300, tol=0.0, printerval=10 ** 9, return_info=True)
    x64, _, i64 = lp.ls_sparse_spectral(y, t, f, ctx=ctx, **kw)
    x32, _, i32 = lp.ls_sparse_spectral(y.astype(np.float32), t, f, ctx=ctx, **kw)
    assert x32.dtype == np.complex64 and x64.dtype == np.complex128
    assert i32["timing"][1] == 0.5 * i64["timing"][1]  # algorithmic bytes per iteration
    # the Float32 signal itself differs from y by 6e-8 relative; the single-precision inverse adds about as much
    y32 = y.astype(np.float32).astype(np.float64)
    xr, _, ir = lp.ls_sparse_spectral(y32, t, f, ctx=ctx, **kw)
    e = rel(i32["z"], ir["z"])
    print(f"float32 ADMM: z vs the double run on the same (rounded) signal {e:.2e}; it/s {300 / i32['timing'][0] * 1e3:.0f} vs "
          f"{300 / ir['timing'][0] * 1e3:.0f}")
    assert e <= 1e-5  # measured 1.9e-6: single-precision entries of a cond-1.15 inverse
    big = np.abs(ir["z"]) > 1e-3 * np.abs(ir["z"]).max()
    assert np.array_equal(i32["z"][big] != 0, ir["z"][big] != 0)
    assert rel(x32, x64.astype(np.complex64)) <= 1e-5
