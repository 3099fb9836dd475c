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
    xr, _ = o.ls_spectral(y, t, f, W, lam=lam, mode="gram") if W is not None else o.ls_spectral(y, t, f, lam=lam,
                                                                                              mode="gram")
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
        Sr, _ = o.ls_windowpsd(y, t, f, nw=nw, noverlap=nov, window_func=o.hanning, mode="gram")
    elif kind == 1:
        S, _ = lp.ls_windowcsd(y, u, t, f, nw=nw, noverlap=nov, window_func=lp.hanning, ctx=ctx)
        Sr, _ = o.ls_windowcsd(y, u, t, f, nw=nw, noverlap=nov, window_func=o.hanning, mode="gram")
    else:
        S, _ = lp.ls_cohere(y, u, t, f, nw=nw, noverlap=nov, ctx=ctx)
        Sr, _ = o.ls_cohere(y, u, t, f, nw=nw, noverlap=nov, mode="gram")
    assert rel(S, Sr) <= 1e-8, (N, n, Nf, nov, kind)
