"""Generates tests/golden/lpvs_golden.npz: seeded inputs + oracle (reference-literal mode) outputs for every entry
point of the hot path.  Julia is not installed, so the vectors come from the oracle, which is itself pinned on the
reference's KATs (tests/test_oracle_kats.py).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lpvs_oracle as o  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    g = {}
    N = 1536
    t = np.sort(10.0 * rng.random(N))
    y = np.sin(2 * np.pi * 11 * t) + 0.6 * np.cos(2 * np.pi * 23.5 * t + 0.7) + 0.1 * rng.standard_normal(N)
    u = 0.7 * np.roll(y, 2) + 0.4 * rng.standard_normal(N)
    f = o.default_freqs(t)[:300]
    g.update(t=t, y=y, u=u, f=f)
    g["ls_unweighted"] = o.ls_spectral(y, t, f, mode="literal")[0]
    W = 0.25 + rng.random(N)
    g["W"] = W
    g["ls_weighted"] = o.ls_spectral(y, t, f[::2], W, mode="literal")[0]
    f_nz = f[1:200]  # no zero frequency, non-multiple of 64
    g["ls_nozero"] = o.ls_spectral(y, t, f_nz, mode="literal")[0]
    f_irr = np.sort(rng.random(70)) * 30 + 0.5  # non-uniform grid -> direct synthesis path
    g["f_irr"] = f_irr
    g["ls_irregular"] = o.ls_spectral(y, t, f_irr, mode="literal")[0]
    nw = 6
    n = N // nw
    fw = np.arange(40) * 2.0 / (t[n] - t[0])
    g["fw"] = fw
    g["psd_hann"] = o.ls_windowpsd(y, t, fw, nw=nw, window_func=o.hanning)[0]
    g["psd_rect_nov0"] = o.ls_windowpsd(y, t, fw[:20], nw=nw, noverlap=0)[0]
    g["csd_hann"] = o.ls_windowcsd(y, u, t, fw, nw=nw, window_func=o.hanning)[0]
    g["cohere"] = o.ls_cohere(y, u, t, fw, nw=nw)[0]
    fs = np.arange(1, 121) * 0.25
    g["fs"] = fs
    x, _, info = o.ls_sparse_spectral(y[:700], t[:700], fs, lam=0.4, iters=1500, tol=1e-9, mode="literal",
                                      return_info=True, printerval=10 ** 9)
    g["sparse_l1"] = x
    g["sparse_l1_iters"] = np.array(info["iters"])
    Y, V, X = o.generate_lpv_signal(400, seed=3)
    w = 2 * np.pi * np.arange(2, 22, 2)
    se = o.ls_spectral_lpv(Y, X, V, w, 16, lam=0.05, mode="literal")
    g.update(lpv_Y=Y, lpv_V=V, lpv_X=X, lpv_w=w, lpv_params=se.x, lpv_sigma_diag=np.diag(se.Sigma).copy(),
             lpv_fva=np.array(se.fva))
    ss, si = o.ls_sparse_spectral_lpv(Y, X, V, w, 16, lam=3.0, iters=1200, tol=1e-8, mode="literal",
                                      return_info=True, printerval=10 ** 9)
    g["sparse_lpv_params"] = ss.x
    g["sparse_lpv_iters"] = np.array(si["iters"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lpvs_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
