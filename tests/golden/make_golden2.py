"""Generates tests/golden/lpvs_golden2.npz: the entry points / options lpvs_golden.npz does not cover -- NormL0 and
IndBallL0 ADMM, the weighted sparse method (sign quirk Q13), init=true (Q14), a windowed estimator with
estimator=ls_sparse_spectral, ls_windowpsd_lpv, and the coulomb / un-normalised LPV bases (dense and sparse, Q16).
Outputs come from the oracle's reference-literal mode (CG x-updates, QR / LU solves).  Julia is not installed, so
these are oracle vectors, not reference vectors; the oracle is pinned by tests/test_oracle_kats.py and
tests/test_oracle_pins.py.  Run:  python tests/golden/make_golden2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lpvs_oracle as o  # noqa: E402

NOPRINT = 10 ** 9


def main():
    rng = np.random.default_rng(20261019)
    g = {}
    N = 900
    t = np.sort(10.0 * rng.random(N))
    y = (np.sin(2 * np.pi * 4.0 * t) + 0.7 * np.cos(2 * np.pi * 9.5 * t + 0.4) + 0.4 * np.sin(2 * np.pi * 17.25 * t)
         + 0.1 * rng.standard_normal(N))
    f = np.arange(0, 100) * 0.25  # zero frequency included -> Nreg = 199
    W = 0.5 + rng.random(N)
    g.update(t=t, y=y, f=f, W=W)
    kw = dict(iters=3000, tol=1e-9, mode="literal", return_info=True, printerval=NOPRINT)
    for name, pg in (("l0", o.NormL0(0.3)), ("ball", o.IndBallL0(7))):
        x, _, info = o.ls_sparse_spectral(y, t, f, proxg=pg, **kw)
        g[f"sparse_{name}"] = x
        g[f"sparse_{name}_iters"] = np.array(info["iters"])
    x, _, info = o.ls_sparse_spectral(y, t, f, W, lam=0.6, **kw)
    g["sparse_weighted"] = x
    g["sparse_weighted_iters"] = np.array(info["iters"])
    x, _, info = o.ls_sparse_spectral(y, t, f, init=True, lam=0.5, **kw)
    g["sparse_init"] = x
    g["sparse_init_iters"] = np.array(info["iters"])
    # windowed estimator with estimator = ls_sparse_spectral (test/test_lasso.jl:36 shape)
    fw = np.arange(1.0, 20.01, 0.5)
    g["fw"] = fw
    est = lambda yi, ti, fr, Wi, **k: o.ls_sparse_spectral(yi, ti, fr, Wi, mode="literal", printerval=NOPRINT, **k)  # noqa: E731
    g["win_sparse_psd"] = o.ls_windowpsd(y, t, fw, nw=3, window_func=o.hanning, estimator=est, lam=0.2, tol=1e-10,
                                         iters=4000, mu=1e-3)[0]
    # LPV family
    Y, V, X = o.generate_lpv_signal(600, seed=7)
    w = 2 * np.pi * np.arange(2, 22, 2)
    g.update(lpv_Y=Y, lpv_V=V, lpv_X=X, lpv_w=w)
    g["windowpsd_lpv"] = o.ls_windowpsd_lpv(Y, X, V, w, 8, nw=3, noverlap=20, lam=0.05, mode="literal")
    Vc = V - 0.5 + 1e-4  # both signs, no sample exactly at 0
    se = o.ls_spectral_lpv(Y, X, Vc, w, 6, lam=0.05, coulomb=True, normalize=False, mode="literal")
    g["lpv_coulomb_params"] = se.x
    g["lpv_coulomb_sigma_diag"] = np.diag(se.Sigma).copy()
    ss, si = o.ls_sparse_spectral_lpv(Y, X, Vc, w, 6, lam=2.0, coulomb=True, iters=1500, tol=1e-8, mode="literal",
                                      return_info=True, printerval=NOPRINT)
    g["sparse_lpv_coulomb_params"] = ss.x
    g["sparse_lpv_coulomb_iters"] = np.array(si["iters"])
    # the quantity the ADMM parity bar is stated on (north_star: objective within 1e-8, identical support)
    g["sparse_lpv_coulomb_objective"] = np.array(o.sparse_objective(si["Phi"], Y, si["z"], si["proxg"]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lpvs_golden2.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")
    for k in sorted(g):
        if k.endswith("_iters"):
            print(k, int(g[k]))


if __name__ == "__main__":
    main()
