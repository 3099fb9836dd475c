"""Seeded random shapes through the dense GPU paths vs the oracle (-m gpu): sizes that are not multiples of any tile,
1..8 column blocks, with / without a zero frequency, weights, one and two right-hand sides, many small windows.
Catches indexing mistakes in the tile / panel / batch logic that fixed-size tests can miss."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(40, 2500))
    Nf = int(rng.integers(1, max(2, min(420, N // 3))))
    zero = bool(rng.integers(0, 2))
    T = 10.0
    t = np.sort(T * rng.random(N))
    df = (0.25 + 0.5 * rng.random()) * N / (4.0 * T * Nf)  # stays below the mean Nyquist: well-conditioned
    f = (np.arange(Nf) + (0 if zero else 1)) * df
    y = sum(rng.standard_normal() * np.cos(2 * np.pi * f[int(rng.integers(0, Nf))] * t + rng.random())
            for _ in range(3)) + 0.1 * rng.standard_normal(N)
    return rng, N, Nf, t, f, y


@pytest.mark.parametrize("seed", range(24))
def test_random_ls_spectral(ctx, seed):
    import lpvspectral_jl_b200 as lp

    rng, N, Nf, t, f, y = _case(seed)
    W = 0.2 + rng.random(N) if seed % 2 else None
    A, _ = o.get_fourier_regressor(t, f)
    G = (A.T * (W if W is not None else 1.0)) @ A
    cond = np.linalg.cond(G)
    if cond > 1e6:
        pytest.skip(f"cond {cond:.1e}")
    lam = 1e-10
    x, _ = lp.ls_spectral(y, t, f, W, lam=lam, ctx=ctx) if W is not None else lp.ls_spectral(y, t, f, lam=lam, ctx=ctx)
    # comparator: the reference's own algorithm (SVD of [A; lam I] unweighted, N-rhs LU weighted), not the Gram restatement
    xr, _ = o.ls_spectral(y, t, f, W, lam=lam, mode="literal") if W is not None else o.ls_spectral(y, t, f, lam=lam,
                                                                                                 mode="literal")
    assert rel(x, xr) <= max(1e-10, 1e-15 * cond), (N, Nf, cond)


@pytest.mark.parametrize("seed", range(12))
def test_random_windowed(ctx, seed):
    import lpvspectral_jl_b200 as lp

    rng, N, Nf, t, f, y = _case(100 + seed)
    N = max(N, 400)
    t = np.sort(10.0 * rng.random(N))
    y = np.sin(2 * np.pi * 3.0 * t) + 0.2 * rng.standard_normal(N)
    u = np.roll(y, 2) + 0.2 * rng.standard_normal(N)
    nw = int(rng.integers(2, 12))
    n = N // nw
    Nf = int(rng.integers(1, max(2, n // 10)))
    f = (np.arange(Nf) + seed % 2) * (2.0 * nw / 10.0)  # spacing 2 / window length: well-conditioned under Hann
    nov = int(rng.integers(0, n - 1)) if seed % 3 else -1
    kind = seed % 3
    if kind == 0:
        S, _ = lp.ls_windowpsd(y, t, f, nw=nw, noverlap=nov, window_func=lp.hanning, ctx=ctx)
        Sr, _ = o.ls_windowpsd(y, t, f, nw=nw, noverlap=nov, window_func=o.hanning, mode="literal")
    elif kind == 1:
        S, _ = lp.ls_windowcsd(y, u, t, f, nw=nw, noverlap=nov, window_func=lp.hanning, ctx=ctx)
        Sr, _ = o.ls_windowcsd(y, u, t, f, nw=nw, noverlap=nov, window_func=o.hanning, mode="literal")
    else:
        S, _ = lp.ls_cohere(y, u, t, f, nw=nw, noverlap=nov, ctx=ctx)
        Sr, _ = o.ls_cohere(y, u, t, f, nw=nw, noverlap=nov, mode="literal")
    assert rel(S, Sr) <= 1e-8, (N, n, Nf, nov, kind)


def support(z):
    return set(np.flatnonzero(np.asarray(z) != 0).tolist())


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("symv", [0, 1])
def test_random_sparse_lpv(ctx, seed, symv):
    """Group lasso over random (Nf, Nv): group sizes from 4 to 60, vectors with and without tile padding, both x-update
    kernels (GEMV over the full inverse / SYMV over its lower triangle with the ADMM-order permutation and whole groups
    per CTA), coulomb on odd seeds (uncovered entries)."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    rng = np.random.default_rng(500 + seed)
    N = int(rng.integers(300, 900))
    Nf = int(rng.integers(2, 14))
    Nv = int(rng.integers(2, 31))
    coul = bool(seed % 2)
    Y, V, X = o.generate_lpv_signal(N, seed=seed)
    if coul:
        V = V - 0.5 + 1e-3  # no sample exactly at 0 (see test_coulomb_zero_sample_is_reported)
    w = 2 * np.pi * np.sort(rng.choice(np.arange(1, 30), size=Nf, replace=False)).astype(float)
    kw = dict(iters=600, tol=1e-8, mu=0.05)
    lam = float(rng.choice([0.5, 2.0, 5.0]))
    ctx.set_option(L.OPT_ADMM_SYMV, symv)
    try:
        se, info = lp.ls_sparse_spectral_lpv(Y, X, V, w, Nv, lam=lam, coulomb=coul, ctx=ctx, return_info=True, **kw)
    finally:
        ctx.set_option(L.OPT_ADMM_SYMV, -1)
    sr, ri = o.ls_sparse_spectral_lpv(Y, X, V, w, Nv, lam=lam, coulomb=coul, mode="gram", return_info=True,
                                      printerval=10 ** 9, **kw)
    assert info["iters"] == ri["iters"], (N, Nf, Nv, coul, lam)
    assert support(info["z"]) == support(ri["z"])
    assert rel(info["z"], ri["z"]) <= 1e-8 if np.linalg.norm(ri["z"]) > 0 else np.all(info["z"] == 0)
    assert rel(info["x"], ri["x"]) <= 1e-8
    if symv == 1:
        # the reference's algorithm proper: warm-started CG x-update (src/lasso.jl:151 -> ProximalOperators / cg!)
        sl, rl = o.ls_sparse_spectral_lpv(Y, X, V, w, Nv, lam=lam, coulomb=coul, mode="literal", return_info=True,
                                          printerval=10 ** 9, **kw)
        assert info["iters"] == rl["iters"], (N, Nf, Nv, coul, lam)
        assert support(info["z"]) == support(rl["z"])
        og = o.sparse_objective(rl["Phi"], Y, info["z"], rl["proxg"])
        ol = o.sparse_objective(rl["Phi"], Y, rl["z"], rl["proxg"])
        assert abs(og - ol) <= 1e-8 * max(1.0, abs(ol))


@pytest.mark.parametrize("seed", range(8))
def test_random_sparse_fourier(ctx, seed):
    """ls_sparse_spectral over random shapes and every Fourier prox operator, weighted (Quadratic, sign quirk) on odd
    seeds: identical iteration count and support, iterates to 1e-8."""
    import lpvspectral_jl_b200 as lp

    rng, N, Nf, t, f, y = _case(200 + seed)
    N = min(N, 900)
    t, y = t[:N], y[:N]
    Nf = min(Nf, 150)
    f = f[:Nf]
    W = 0.3 + rng.random(N) if seed % 2 else None
    pg, pgo = [(lp.NormL1(0.3), o.NormL1(0.3)), (lp.NormL0(0.05), o.NormL0(0.05)),
               (lp.IndBallL0(5), o.IndBallL0(5))][seed % 3]
    kw = dict(iters=500, tol=1e-9, mu=0.05)
    x, _, info = lp.ls_sparse_spectral(y, t, f, W, proxg=pg, ctx=ctx, return_info=True, **kw)
    xr, _, ri = o.ls_sparse_spectral(y, t, f, W, proxg=pgo, mode="gram", return_info=True, printerval=10 ** 9, **kw)
    assert info["iters"] == ri["iters"], (N, Nf, seed)
    assert support(info["z"]) == support(ri["z"])
    assert rel(info["z"], ri["z"]) <= 1e-8
    # and against the reference's algorithm proper (warm-started CG x-update): iteration count, support, objective
    xl, _, rl = o.ls_sparse_spectral(y, t, f, W, proxg=pgo, mode="literal", return_info=True, printerval=10 ** 9, **kw)
    assert info["iters"] == rl["iters"], (N, Nf, seed)
    assert support(info["z"]) == support(rl["z"])
    ysign = -y if W is not None else y  # the weighted method fits -y (Q13); the objective is evaluated accordingly
    Aw = rl["A"] * np.sqrt(W)[:, None] if W is not None else rl["A"]
    yw = ysign * np.sqrt(W) if W is not None else ysign
    og, ol = o.sparse_objective(Aw, yw, info["z"], pgo), o.sparse_objective(Aw, yw, rl["z"], pgo)
    assert abs(og - ol) <= 1e-8 * max(1.0, abs(ol))


def test_coulomb_zero_sample_is_reported(ctx):
    """coulomb=true masks every basis function for a scheduling value whose sign matches no centre (V == 0), so the
    reference's normalisation K/sum(K) is 0/0 and it returns NaNs (src/lsfft.jl:199-207).  The library reports the cause
    (LPVS_E_NONFINITE) instead of a misleading factorisation failure; found by the random-shape tests."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    Y, V, X = o.generate_lpv_signal(401, seed=1)
    V = V - 0.5
    assert np.any(V == 0.0)
    w = 2 * np.pi * np.array([2.0, 10.0])
    with pytest.raises(lp.LpvsError) as e1:
        lp.ls_sparse_spectral_lpv(Y, X, V, w, 6, lam=1.0, coulomb=True, iters=50, ctx=ctx)
    assert e1.value.code == L.E_NONFINITE and "0/0" in str(e1.value)
    with pytest.raises(lp.LpvsError) as e2:
        lp.ls_spectral_lpv(Y, X, V, w, 6, lam=0.05, coulomb=True, ctx=ctx)
    assert e2.value.code == L.E_NONFINITE
    # the oracle (like the reference) just propagates NaN
    sr = o.ls_spectral_lpv(Y, X, V, w, 6, lam=0.05, coulomb=True, mode="gram")
    assert not np.all(np.isfinite(sr.x))
    # and without the offending sample everything is fine
    keep = V != 0.0
    se = lp.ls_spectral_lpv(Y[keep], X[keep], V[keep], w, 6, lam=0.05, coulomb=True, ctx=ctx)
    so = o.ls_spectral_lpv(Y[keep], X[keep], V[keep], w, 6, lam=0.05, coulomb=True, mode="gram")
    assert rel(se.x, so.x) <= 1e-9


@pytest.mark.parametrize("seed", range(10))
def test_random_ls_spectral_lpv(ctx, seed):
    """Dense LPV estimator over random (N, Nf, Nv), all four coulomb / normalize combinations: parameters, covariance
    and fraction of variance explained vs the oracle."""
    import warnings

    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(900 + seed)
    N = int(rng.integers(200, 1500))
    Nf = int(rng.integers(1, 12))
    Nv = int(rng.integers(2, 25))
    coul, norm = bool(seed & 1), bool(seed & 2)
    Y, V, X = o.generate_lpv_signal(N, seed=seed)
    if coul:
        V = V - 0.5 + 1e-3
    w = 2 * np.pi * np.sort(rng.choice(np.arange(1, 25), size=Nf, replace=False)).astype(float)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        se = lp.ls_spectral_lpv(Y, X, V, w, Nv, lam=0.05, coulomb=coul, normalize=norm, ctx=ctx)
        so = o.ls_spectral_lpv(Y, X, V, w, Nv, lam=0.05, coulomb=coul, normalize=norm, mode="literal")
    assert se.x.shape == so.x.shape
    assert rel(se.x, so.x) <= 1e-8, (N, Nf, Nv, coul, norm)
    assert rel(se.Σ, so.Sigma) <= 1e-8
    assert abs(se.fva - so.fva) <= 1e-9
    assert rel(lp.psd(se), o.psd(so)) <= 1e-8
