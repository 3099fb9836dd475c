"""Second golden set (tests/golden/lpvs_golden2.npz, made by tests/golden/make_golden2.py from the oracle's
reference-literal mode): NormL0 / IndBallL0 ADMM, the weighted sparse method (Q13), init=true (Q14), a windowed
estimator with estimator=ls_sparse_spectral, ls_windowpsd_lpv, coulomb / un-normalised LPV bases (dense, sparse Q16).
CPU: the oracle's Gram mode (what the GPU computes) reproduces them; GPU: liblpvs reproduces them."""
import os

import numpy as np
import pytest

from oracle import lpvs_oracle as o

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpvs_golden2.npz"))
TOL = 1e-9       # coefficients of direct solves (north_star)
TOL_ADMM = 1e-8  # ADMM results against the CG-based literal run (north_star: objective 1e-8, identical support)
NOPRINT = 10 ** 9


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def support(x):
    return set(np.flatnonzero(np.asarray(x)).tolist())


def lpv_objective(x, Y, X, V, w, Nv, lam, coulomb):
    """0.5||Phi z - Y||^2 + lam sum_f ||z_group f||_2 of a sparse-LPV result x (src/lasso.jl:44-55)."""
    Ar = o.lpv_regressor(X, V, w, Nv, True, coulomb)
    inds = o.lpv_group_perm(len(w), Ar.shape[1])
    z = np.concatenate([x.real, x.imag])[inds]
    return o.sparse_objective(Ar[:, inds], Y, z, o.GroupNormL2(lam, 2 * Nv, len(w)))


def check_sparse_lpv_coulomb(x, iters, Y, X, Vc, w):
    # at tol 1e-8 the CG-based literal run and an exact x-update agree to ~1e-7 in the coefficients; the bar is the
    # objective (1e-8 relative) with an identical support and iteration count
    assert iters == int(G["sparse_lpv_coulomb_iters"])
    assert support(x) == support(G["sparse_lpv_coulomb_params"])
    obj = lpv_objective(x, Y, X, Vc, w, 6, 2.0, True)
    assert abs(obj - float(G["sparse_lpv_coulomb_objective"])) <= 1e-8 * float(G["sparse_lpv_coulomb_objective"])
    assert rel(x, G["sparse_lpv_coulomb_params"]) <= 1e-6


def test_oracle_gram_mode_reproduces_golden2():
    t, y, f, W = G["t"], G["y"], G["f"], G["W"]
    kw = dict(iters=3000, tol=1e-9, mode="gram", return_info=True, printerval=NOPRINT)
    for name, pg in (("l0", o.NormL0(0.3)), ("ball", o.IndBallL0(7))):
        x, _, info = o.ls_sparse_spectral(y, t, f, proxg=pg, **kw)
        assert info["iters"] == int(G[f"sparse_{name}_iters"]) and support(x) == support(G[f"sparse_{name}"])
        assert rel(x, G[f"sparse_{name}"]) <= TOL_ADMM
    x, _, info = o.ls_sparse_spectral(y, t, f, W, lam=0.6, **kw)
    assert info["iters"] == int(G["sparse_weighted_iters"]) and support(x) == support(G["sparse_weighted"])
    assert rel(x, G["sparse_weighted"]) <= TOL_ADMM
    x, _, info = o.ls_sparse_spectral(y, t, f, init=True, lam=0.5, **kw)
    assert info["iters"] == int(G["sparse_init_iters"]) and rel(x, G["sparse_init"]) <= TOL_ADMM
    est = lambda yi, ti, fr, Wi, **k: o.ls_sparse_spectral(yi, ti, fr, Wi, mode="gram", printerval=NOPRINT, **k)  # noqa: E731
    S = o.ls_windowpsd(y, t, G["fw"], nw=3, window_func=o.hanning, estimator=est, lam=0.2, tol=1e-10, iters=4000,
                       mu=1e-3)[0]
    assert rel(S, G["win_sparse_psd"]) <= TOL_ADMM
    Y, V, X, w = G["lpv_Y"], G["lpv_V"], G["lpv_X"], G["lpv_w"]
    assert rel(o.ls_windowpsd_lpv(Y, X, V, w, 8, nw=3, noverlap=20, lam=0.05, mode="gram"), G["windowpsd_lpv"]) <= TOL
    Vc = V - 0.5 + 1e-4
    se = o.ls_spectral_lpv(Y, X, Vc, w, 6, lam=0.05, coulomb=True, normalize=False, mode="gram")
    assert rel(se.x, G["lpv_coulomb_params"]) <= TOL
    assert rel(np.diag(se.Sigma), G["lpv_coulomb_sigma_diag"]) <= TOL
    ss, si = o.ls_sparse_spectral_lpv(Y, X, Vc, w, 6, lam=2.0, coulomb=True, iters=1500, tol=1e-8, mode="gram",
                                      return_info=True, printerval=NOPRINT)
    check_sparse_lpv_coulomb(ss.x, si["iters"], Y, X, Vc, w)


@pytest.mark.gpu
def test_gpu_reproduces_golden2(ctx):
    import lpvspectral_jl_b200 as lp

    t, y, f, W = G["t"], G["y"], G["f"], G["W"]
    kw = dict(iters=3000, tol=1e-9, ctx=ctx, return_info=True)
    for name, pg in (("l0", lp.NormL0(0.3)), ("ball", lp.IndBallL0(7))):
        x, _, info = lp.ls_sparse_spectral(y, t, f, proxg=pg, **kw)
        assert info["iters"] == int(G[f"sparse_{name}_iters"]) and support(x) == support(G[f"sparse_{name}"])
        assert rel(x, G[f"sparse_{name}"]) <= TOL_ADMM
    x, _, info = lp.ls_sparse_spectral(y, t, f, W, lam=0.6, **kw)
    assert info["iters"] == int(G["sparse_weighted_iters"]) and support(x) == support(G["sparse_weighted"])
    assert rel(x, G["sparse_weighted"]) <= TOL_ADMM
    x, _, info = lp.ls_sparse_spectral(y, t, f, init=True, lam=0.5, **kw)
    assert info["iters"] == int(G["sparse_init_iters"]) and rel(x, G["sparse_init"]) <= TOL_ADMM
    S = lp.ls_windowpsd(y, t, G["fw"], nw=3, window_func=lp.hanning, estimator=lp.ls_sparse_spectral, lam=0.2,
                        tol=1e-10, iters=4000, mu=1e-3, ctx=ctx)[0]
    assert rel(S, G["win_sparse_psd"]) <= TOL_ADMM
    Y, V, X, w = G["lpv_Y"], G["lpv_V"], G["lpv_X"], G["lpv_w"]
    assert rel(lp.ls_windowpsd_lpv(Y, X, V, w, 8, 3, 20, lam=0.05, ctx=ctx), G["windowpsd_lpv"]) <= TOL
    Vc = V - 0.5 + 1e-4
    se = lp.ls_spectral_lpv(Y, X, Vc, w, 6, lam=0.05, coulomb=True, normalize=False, ctx=ctx)
    assert rel(se.x, G["lpv_coulomb_params"]) <= TOL
    assert rel(np.diag(se.Σ), G["lpv_coulomb_sigma_diag"]) <= TOL
    ss, si = lp.ls_sparse_spectral_lpv(Y, X, Vc, w, 6, lam=2.0, coulomb=True, iters=1500, tol=1e-8, ctx=ctx,
                                       return_info=True)
    check_sparse_lpv_coulomb(ss.x, si["iters"], Y, X, Vc, w)
