"""LPVS_PHASE_STRUCTURED (opt-in): Gram matrices of a uniform-grid Fourier basis from their 3 Nf trigonometric sums
(csrc/structured.cu: Toeplitz + Hankel in the frequency index) against the oracle, through every caller of the Gram stage.

The mode carries the exact phase of the ideal grid f0 + k df -- the accuracy class of LPVS_PHASE_CHAIN -- so on ordinary
phases (|2 pi f t| << 1/sqrt(eps)) it meets the same 1e-9 bar against the reference-literal oracle as the default mode."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(params=["structured", "structured_ref"])
def sctx(ctx, request):
    """Both opt-in modes: the plain sums, and the sums + the first-order correction for the reference's phase rounding
    (LPVS_PHASE_STRUCTURED_REF, csrc/corr.cu) -- on ordinary phases the correction is far below every bar used here, so the
    same tests exercise its plumbing (scales, padding columns, ragged chunks, sample splits, two right-hand sides)."""
    from lpvspectral_jl_b200 import _lib as L

    ctx.set_option(L.OPT_PHASE_MODE, L.PHASE_STRUCTURED if request.param == "structured" else L.PHASE_STRUCTURED_REF)
    yield ctx
    ctx.set_option(L.OPT_PHASE_MODE, L.PHASE_AUTO)


@pytest.mark.parametrize("N,Nf,f0,df", [(5000, 64, 0.0, 0.31), (4097, 65, 0.4, 0.2), (9000, 200, 0.0, 0.11),
                                         (3000, 1, 0.0, 0.5), (2100, 7, 1.3, 0.9), (70000, 130, 0.05, 0.05)])
def test_gram_from_sums_matches_the_product_form(sctx, N, Nf, f0, df):
    """G and b of one problem (sample splits and all) vs A'WA, A'Wy of the oracle's regressor, weighted and not."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(N + Nf)
    t = np.sort(10 * rng.random(N))
    y = rng.standard_normal(N)
    f = f0 + df * np.arange(Nf)
    A, _ = o.get_fourier_regressor(t, f)
    for W in (None, 0.5 + rng.random(N)):
        G, b = lp.gram_fourier(t, f, W, y, ctx=sctx)
        Aw = A if W is None else A * W[:, None]
        Gr, br = Aw.T @ A, Aw.T @ y
        assert np.abs(G - Gr).max() <= 5e-13 * np.abs(Gr).max()
        assert np.abs(b - br).max() <= 5e-13 * np.abs(br).max()
        assert np.array_equal(G, G.T)


def test_non_uniform_grid_is_rejected(sctx):
    import lpvspectral_jl_b200 as lp

    t = np.sort(np.random.default_rng(0).random(500))
    with pytest.raises(Exception, match="uniform"):
        lp.gram_fourier(t, np.array([0.0, 1.0, 2.5, 3.0]), ctx=sctx)


def test_estimators_through_the_structured_gram(sctx):
    """ls_spectral (weighted / unweighted), windowed psd / coherence and the L1 ADMM against the oracle's literal modes."""
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(7)
    N = 6000
    t = np.sort(10 * rng.random(N))
    y = np.sin(2 * np.pi * 20 * t) + 0.5 * np.cos(2 * np.pi * 55 * t + 1) + 0.1 * rng.standard_normal(N)
    u = 0.7 * np.roll(y, 3) + 0.3 * rng.standard_normal(N)
    f = o.default_freqs(t)[:700]
    x, _ = lp.ls_spectral(y, t, f, ctx=sctx)
    xr, _ = o.ls_spectral(y, t, f, mode="literal")
    assert rel(x, xr) <= 1e-9
    W = 0.5 + rng.random(N)
    xw, _ = lp.ls_spectral(y, t, f, W, ctx=sctx)
    xwr, _ = o.ls_spectral(y, t, f, W, mode="literal")
    assert rel(xw, xwr) <= 1e-9
    n = N // 12
    fw = np.arange(40) * 2.0 / (t[n] - t[0])
    S, _ = lp.ls_windowpsd(y, t, fw, nw=12, window_func=lp.hanning, ctx=sctx)
    Sr, _ = o.ls_windowpsd(y, t, fw, nw=12, window_func=o.hanning)
    assert rel(S, Sr) <= 1e-9
    Cxy, _ = lp.ls_cohere(y, u, t, fw, nw=12, ctx=sctx)
    Cr, _ = o.ls_cohere(y, u, t, fw, nw=12)
    assert rel(Cxy, Cr) <= 1e-9
    Cyy, _ = lp.ls_cohere(y, y, t, fw, nw=12, ctx=sctx)
    assert np.all(Cyy == 1)
    # windowed estimator with estimator = ls_sparse_spectral (batched ADMM on the windows' Gram matrices)
    kw = dict(tol=1e-9, iters=1500, mu=0.05)
    est = lambda yi, ti, fr, W, **k: o.ls_sparse_spectral(yi, ti, fr, W, mode="gram", printerval=10 ** 9, **k)  # noqa: E731
    Cs, _ = lp.ls_cohere(y, u, t, fw, nw=12, estimator=lp.ls_sparse_spectral, proxg=lp.NormL1(0.05), ctx=sctx, **kw)
    Csr, _ = o.ls_cohere(y, u, t, fw, nw=12, estimator=est, proxg=o.NormL1(0.05), **kw)
    ok = np.isfinite(Csr)
    assert np.array_equal(ok, np.isfinite(Cs)) and rel(Cs[ok], Csr[ok]) <= 1e-8
    fs = np.arange(1, 129) * 0.5
    z, _ = lp.ls_sparse_spectral(y[:1024], t[:1024], fs, lam=0.2, iters=400, tol=1e-9, ctx=sctx)
    zr, _ = o.ls_sparse_spectral(y[:1024], t[:1024], fs, lam=0.2, iters=400, tol=1e-9, printerval=10 ** 9)
    assert rel(z, zr) <= 1e-8 and np.array_equal(z != 0, zr != 0)


def test_cfg2_windows_meet_the_parity_bar(sctx):
    """BASELINE configs[1] at full size in the structured mode: sampled windows against the literal N-rhs LU.  At cfg2's phases
    (<= 3.3e6 rad) the reference's own phase rounding is worth <= 4e-10 rad per element, so the exact-phase class still meets
    the flat 1e-9 bar on well-conditioned windows (it does not at cfg5a's 2.6e7 rad: DESIGN.md section 1b)."""
    import bench
    import scipy.linalg as sla
    from lpvspectral_jl_b200 import _lib as L
    import lpvspectral_jl_b200 as lp

    t, y, f, n = bench.make_cfg2()
    hop = n >> 1
    W = o.hanning(n)
    rng = np.random.default_rng(3)
    worst = 0.0
    for k in sorted(set([0, 2046] + list(rng.integers(0, 2047, 10)))):
        sl = slice(k * hop, k * hop + n)
        s = lp.window_sums(L.WIN_PSD, y, None, t, f, W, n, hop, 1e-10, k, k + 1, ctx=sctx)
        A, zf = o.get_fourier_regressor(t[sl], f)
        AtW = A.T * W
        M = AtW @ A + 1e-10 * np.eye(A.shape[1])
        cond = np.linalg.cond(M)
        X = sla.lu_solve(sla.lu_factor(M, check_finite=False), AtW, check_finite=False)
        ref = o._abs2(o.fourier2complex(X @ y[sl], zf))
        e = rel(s, ref)
        worst = max(worst, e / max(1e-9, 40.0 * cond * 2.2e-16))
        assert e <= max(1e-9, 40.0 * cond * 2.2e-16), (k, cond, e)
    # the whole record: linear in y^2 and equal to the default mode's answer to the same bar
    S, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=sctx)
    sctx.set_option(L.OPT_PHASE_MODE, L.PHASE_AUTO)
    Sd, _ = lp.ls_windowpsd(y, t, f, nw=1024, window_func=lp.hanning, ctx=sctx)
    print(f"cfg2 structured: worst sampled-window error / bar {worst:.3f}; whole PSD vs the default mode {rel(S, Sd):.2e}")
    assert rel(S, Sd) <= 1e-9


@pytest.mark.parametrize("weighted", [False, True])
def test_structured_ref_reproduces_reference_phase_rounding(ctx, weighted):
    """The setting of test_chain_ref_reproduces_reference_phase_rounding (phi ~ 2.6e7 rad: the reference's fl(fl(2 pi f) t) is off
    the ideal phase by up to 2.9e-9 rad).  The plain structured Gram matrix differs from the oracle's by about that; with the
    first-order correction G += D'B + B'D, b += D'y (half-precision tensor-core GEMM, eps exact in FP64) what is left is the
    f16 operand rounding: <= 1e-3 of the difference (measured 3e-4), i.e. the class of the per-element modes."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    rng = np.random.default_rng(12)
    N, Nf = 4096, 200
    t = 9.99 + np.sort(2.4e-3 * rng.random(N))
    fs = N / 2.4e-3
    f = np.arange(Nf) * (fs / N) * 5.0
    W = 3.0 * o.hanning(N) if weighted else None
    y = rng.standard_normal(N)
    A, _ = o.get_fourier_regressor(t, f)
    Aw = A if W is None else A * W[:, None]
    Gr, br = Aw.T @ A, Aw.T @ y
    err = {}
    for mode in (L.PHASE_STRUCTURED, L.PHASE_STRUCTURED_REF, L.PHASE_DIRECT):
        ctx.set_option(L.OPT_PHASE_MODE, mode)
        try:
            G, b = lp.gram_fourier(t, f, W, y, ctx=ctx)
        finally:
            ctx.set_option(L.OPT_PHASE_MODE, 0)
        err[mode] = (np.abs(G - Gr).max() / np.abs(Gr).max(), np.abs(b - br).max() / np.abs(br).max())
    print(f"Gram / rhs error vs the oracle (reference rounding): structured {err[L.PHASE_STRUCTURED][0]:.1e} / "
          f"{err[L.PHASE_STRUCTURED][1]:.1e}, structured_ref {err[L.PHASE_STRUCTURED_REF][0]:.1e} / "
          f"{err[L.PHASE_STRUCTURED_REF][1]:.1e}, direct {err[L.PHASE_DIRECT][0]:.1e} / {err[L.PHASE_DIRECT][1]:.1e}")
    assert err[L.PHASE_STRUCTURED][0] > 1e-11  # the effect is there to be corrected
    for q in (0, 1):
        assert err[L.PHASE_STRUCTURED_REF][q] <= 5e-13 + 2e-3 * err[L.PHASE_STRUCTURED][q]


def test_structured_ref_large_problem_with_sample_splits(ctx):
    """One tall problem (sample splits -> partial buffers, added in split order) at large phases, two right-hand sides through
    ls_spectral's weighted path: structured_ref vs the per-element reference-phase mode."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    rng = np.random.default_rng(21)
    N, Nf = 300000, 100
    t = 5.0e4 + np.sort(50.0 * rng.random(N))
    f = np.arange(Nf) * 0.37
    W = 0.5 + rng.random(N)
    y = np.sin(2 * np.pi * 3.7 * t) + 0.1 * rng.standard_normal(N)
    out = {}
    for mode in (L.PHASE_STRUCTURED, L.PHASE_STRUCTURED_REF, L.PHASE_DIRECT):
        ctx.set_option(L.OPT_PHASE_MODE, mode)
        try:
            out[mode] = lp.gram_fourier(t, f, W, y, ctx=ctx)
        finally:
            ctx.set_option(L.OPT_PHASE_MODE, 0)
    Gd, bd = out[L.PHASE_DIRECT]
    e0 = np.abs(out[L.PHASE_STRUCTURED][0] - Gd).max() / np.abs(Gd).max()
    e1 = np.abs(out[L.PHASE_STRUCTURED_REF][0] - Gd).max() / np.abs(Gd).max()
    b0 = np.abs(out[L.PHASE_STRUCTURED][1] - bd).max() / np.abs(bd).max()
    b1 = np.abs(out[L.PHASE_STRUCTURED_REF][1] - bd).max() / np.abs(bd).max()
    print(f"tall problem: G structured {e0:.1e} -> structured_ref {e1:.1e}; b {b0:.1e} -> {b1:.1e} (vs direct mode)")
    assert e1 <= 5e-13 + 2e-3 * e0 and b1 <= 5e-13 + 2e-3 * b0
    assert np.array_equal(out[L.PHASE_STRUCTURED_REF][0], out[L.PHASE_STRUCTURED_REF][0].T)


def test_structured_ref_table_segments(ctx):
    """A problem whose correction tables exceed one 4 GiB segment (2048 frequencies: 6 KB of table per sample, 800 k samples ->
    two segments, each with its own power-of-two scales, accumulating into the same sample-split partials): structured_ref vs
    the per-element reference-phase mode at large phases."""
    import lpvspectral_jl_b200 as lp
    from lpvspectral_jl_b200 import _lib as L

    rng = np.random.default_rng(33)
    N, Nf = 800_000, 2048
    t = 2.0e4 + np.sort(40.0 * rng.random(N))
    f = np.arange(Nf) * 0.21
    W = 0.5 + rng.random(N)
    y = np.sin(2 * np.pi * 37.8 * t) + 0.1 * rng.standard_normal(N)
    out = {}
    for mode in (L.PHASE_STRUCTURED, L.PHASE_STRUCTURED_REF, L.PHASE_DIRECT):
        ctx.set_option(L.OPT_PHASE_MODE, mode)
        try:
            out[mode] = lp.gram_fourier(t, f, W, y, ctx=ctx)
        finally:
            ctx.set_option(L.OPT_PHASE_MODE, 0)
    Gd, bd = out[L.PHASE_DIRECT]
    e0 = np.abs(out[L.PHASE_STRUCTURED][0] - Gd).max() / np.abs(Gd).max()
    e1 = np.abs(out[L.PHASE_STRUCTURED_REF][0] - Gd).max() / np.abs(Gd).max()
    b0 = np.abs(out[L.PHASE_STRUCTURED][1] - bd).max() / np.abs(bd).max()
    b1 = np.abs(out[L.PHASE_STRUCTURED_REF][1] - bd).max() / np.abs(bd).max()
    print(f"two table segments: G structured {e0:.1e} -> structured_ref {e1:.1e}; b {b0:.1e} -> {b1:.1e} (vs direct mode)")
    assert e1 <= 5e-13 + 2e-3 * e0 and b1 <= 5e-13 + 2e-3 * b0
    ctx.release_workspace()


def test_unknown_phase_mode_is_rejected(ctx):
    """lpvs_set_option validates LPVS_OPT_PHASE_MODE (an unknown value used to fall through to the per-element mode silently)."""
    from lpvspectral_jl_b200 import _lib as L

    for bad in (6, -1, 99):
        with pytest.raises(Exception, match="PHASE_MODE"):
            ctx.set_option(L.OPT_PHASE_MODE, bad)
    ctx.set_option(L.OPT_PHASE_MODE, L.PHASE_AUTO)
