"""Golden vectors (tests/golden/lpvs_golden.npz, made by tests/golden/make_golden.py from the oracle's
reference-literal mode).  CPU: the oracle's Gram mode reproduces them; GPU: liblpvs reproduces them."""
import os

import numpy as np
import pytest

from oracle import lpvs_oracle as o

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpvs_golden.npz"))
TOL = 1e-9


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_oracle_gram_mode_reproduces_golden():
    t, y, f = G["t"], G["y"], G["f"]
    assert rel(o.ls_spectral(y, t, f, mode="gram")[0], G["ls_unweighted"]) <= TOL
    assert rel(o.ls_spectral(y, t, f[::2], G["W"], mode="gram")[0], G["ls_weighted"]) <= TOL
    assert rel(o.ls_windowpsd(y, t, G["fw"], nw=6, window_func=o.hanning, mode="gram")[0], G["psd_hann"]) <= TOL
    x, _, info = o.ls_sparse_spectral(y[:700], t[:700], G["fs"], lam=0.4, iters=1500, tol=1e-9, mode="gram",
                                      return_info=True, printerval=10 ** 9)
    assert info["iters"] == int(G["sparse_l1_iters"])
    assert rel(x, G["sparse_l1"]) <= 1e-8
    se = o.ls_spectral_lpv(G["lpv_Y"], G["lpv_X"], G["lpv_V"], G["lpv_w"], 16, lam=0.05, mode="gram")
    assert rel(se.x, G["lpv_params"]) <= TOL


@pytest.mark.gpu
def test_gpu_reproduces_golden(ctx):
    import lpvspectral_jl_b200 as lp

    t, y, u, f = G["t"], G["y"], G["u"], G["f"]
    assert rel(lp.ls_spectral(y, t, f, ctx=ctx)[0], G["ls_unweighted"]) <= TOL
    assert rel(lp.ls_spectral(y, t, f[::2], G["W"], ctx=ctx)[0], G["ls_weighted"]) <= TOL
    assert rel(lp.ls_spectral(y, t, f[1:200], ctx=ctx)[0], G["ls_nozero"]) <= TOL
    assert rel(lp.ls_spectral(y, t, G["f_irr"], ctx=ctx)[0], G["ls_irregular"]) <= TOL
    fw = G["fw"]
    assert rel(lp.ls_windowpsd(y, t, fw, nw=6, window_func=lp.hanning, ctx=ctx)[0], G["psd_hann"]) <= TOL
    assert rel(lp.ls_windowpsd(y, t, fw[:20], nw=6, noverlap=0, ctx=ctx)[0], G["psd_rect_nov0"]) <= TOL
    assert rel(lp.ls_windowcsd(y, u, t, fw, nw=6, window_func=lp.hanning, ctx=ctx)[0], G["csd_hann"]) <= TOL
    assert rel(lp.ls_cohere(y, u, t, fw, nw=6, ctx=ctx)[0], G["cohere"]) <= TOL
    x, _, info = lp.ls_sparse_spectral(y[:700], t[:700], G["fs"], lam=0.4, iters=1500, tol=1e-9, ctx=ctx,
                                       return_info=True)
    assert info["iters"] == int(G["sparse_l1_iters"])
    assert set(np.flatnonzero(x)) == set(np.flatnonzero(G["sparse_l1"]))
    assert rel(x, G["sparse_l1"]) <= 1e-8
    se = lp.ls_spectral_lpv(G["lpv_Y"], G["lpv_X"], G["lpv_V"], G["lpv_w"], 16, lam=0.05, ctx=ctx)
    assert rel(se.x, G["lpv_params"]) <= TOL
    assert rel(np.diag(se.Σ), G["lpv_sigma_diag"]) <= TOL
    assert abs(se.fva - float(G["lpv_fva"])) <= 1e-10
    ss, si = lp.ls_sparse_spectral_lpv(G["lpv_Y"], G["lpv_X"], G["lpv_V"], G["lpv_w"], 16, lam=3.0, iters=1200,
                                       tol=1e-8, ctx=ctx, return_info=True)
    assert si["iters"] == int(G["sparse_lpv_iters"])
    assert rel(ss.x, G["sparse_lpv_params"]) <= 1e-8
