"""Independent pins for the parts of the oracle the reference's own tests do not pin ("parity
unpinned"): the ADMM sub-steps live in ProximalOperators.jl / IterativeSolvers.jl and test/test_lasso.jl has no
assertions.  These tests check the oracle's restatement of those semantics against things that do NOT share code
with it: the defining variational problem of every proximal operator (brute force), the optimality (KKT)
conditions of the problems ADMM converges to, scikit-learn's coordinate-descent lasso, scipy's CG, an mpmath
evaluation of the regressor, and a brute-force restatement of DSP.arraysplit.  CPU only."""
import math

import numpy as np
import pytest

from oracle import lpvs_oracle as o


# ---------------------------------------------------------------------------------------------------------------
# proximal operators: prox_{γg}(v) = argmin_x g(x) + ‖x−v‖²/(2γ)   (ProximalOperators.jl's definition of prox!)
# ---------------------------------------------------------------------------------------------------------------


def _moreau(g, x, v, gamma):
    d = x - v
    return g.value(x) + float(d @ d) / (2.0 * gamma)


@pytest.mark.parametrize("lam,gamma", [(0.1, 0.05), (1.0, 0.05), (0.4, 1.0)])
def test_prox_l1_l0_minimise_their_moreau_objective_scalar(lam, gamma):
    # separable penalties: a dense scalar grid (that contains 0 and v) must not beat the closed form
    grid = np.linspace(-3.0, 3.0, 6001)
    for g in (o.NormL1(lam), o.NormL0(lam)):
        for v in np.linspace(-2.5, 2.5, 41):
            p = g.prox(np.array([v]), gamma)
            best = min(_moreau(g, np.array([x]), np.array([v]), gamma) for x in np.concatenate([grid, [0.0, v]]))
            assert _moreau(g, p, np.array([v]), gamma) <= best + 1e-12


def test_prox_l0_threshold_is_sqrt_2_gamma_lambda():
    g, gamma = o.NormL0(0.3), 0.05
    thr = math.sqrt(2 * gamma * 0.3)
    v = np.array([np.nextafter(thr, 0), thr, np.nextafter(thr, 1), -np.nextafter(thr, 1), 0.0])
    assert np.array_equal(g.prox(v, gamma), np.array([0.0, 0.0, v[2], v[3], 0.0]))  # strict '>' keeps


def test_prox_indball_l0_is_the_projection_and_breaks_ties_low_index():
    rng = np.random.default_rng(0)
    v = rng.standard_normal(9)
    for r in (1, 3, 9):
        p = o.IndBallL0(r).prox(v, 0.05)
        assert np.count_nonzero(p) == r
        # brute force over all supports of size r: the projection keeps the r largest magnitudes
        from itertools import combinations

        best = min(combinations(range(9), r), key=lambda s: np.sum(np.delete(v, list(s)) ** 2))
        assert set(np.flatnonzero(p)) == set(best)
    tie = np.array([1.0, -2.0, 2.0, 0.5, 2.0])
    assert np.array_equal(np.flatnonzero(o.IndBallL0(2).prox(tie, 1.0)), [1, 2])


def test_prox_group_l2_is_block_soft_threshold():
    rng = np.random.default_rng(1)
    g = o.GroupNormL2(0.7, 4, 3)
    v = rng.standard_normal(12) * np.repeat([0.01, 1.0, 3.0], 4)
    gamma = 0.5
    p = g.prox(v, gamma)
    assert np.all(p[:4] == 0)  # ‖v_g‖ < γλ: the whole group dies
    # first-order condition of the Moreau objective on the surviving groups: (p−v)/γ + λ p/‖p‖ = 0
    for k in (1, 2):
        pg, vg = p[4 * k:4 * k + 4], v[4 * k:4 * k + 4]
        assert np.allclose((pg - vg) / gamma + 0.7 * pg / np.linalg.norm(pg), 0, atol=1e-13)
    # and random perturbations never improve it
    base = _moreau(g, p, v, gamma)
    for _ in range(200):
        assert _moreau(g, p + 1e-3 * rng.standard_normal(12), v, gamma) >= base
    # Q16: entries not covered by a group are zeroed
    g2 = o.GroupNormL2(0.7, 4, 2)
    assert np.all(g2.prox(v, gamma)[8:] == 0)


# ---------------------------------------------------------------------------------------------------------------
# x-update: the warm-started CG restatement against scipy's CG and against the exact solve
# ---------------------------------------------------------------------------------------------------------------


def test_cg_restatement_matches_scipy_and_exact_solve():
    from scipy.sparse.linalg import LinearOperator, cg

    rng = np.random.default_rng(2)
    B = rng.standard_normal((60, 40))
    G, mu = B.T @ B, 0.05
    M = G + np.eye(40) / mu
    rhs, x0 = rng.standard_normal(40), rng.standard_normal(40)
    x, its = o._cg(lambda p: M @ p, rhs, x0, math.sqrt(np.finfo(float).eps), 40)
    exact = np.linalg.solve(M, rhs)
    r0 = np.linalg.norm(rhs - M @ x0)
    # stops at the first iterate whose residual is ≤ √eps · the INITIAL residual (warm start matters)
    assert np.linalg.norm(rhs - M @ x) <= 1.0001 * math.sqrt(np.finfo(float).eps) * r0 and 0 < its <= 40
    assert np.linalg.norm(x - exact) <= 1e-5 * np.linalg.norm(exact)
    xs, info = cg(LinearOperator((40, 40), matvec=lambda p: M @ p), rhs, x0=x0, rtol=0.0,
                  atol=math.sqrt(np.finfo(float).eps) * r0, maxiter=40)
    assert info == 0 and np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(exact)
    # the two QuadProx modes are the same operator
    for kind in ("ls", "quad"):
        a = o.QuadProx(G, rhs, kind, "literal").prox(x0, mu)
        b = o.QuadProx(G, rhs, kind, "gram").prox(x0, mu)
        assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(b)


# ---------------------------------------------------------------------------------------------------------------
# ADMM fixed points: optimality conditions of the problems the reference states (src/lasso.jl:77-84)
# ---------------------------------------------------------------------------------------------------------------


def _fourier_problem(N=300, Nf=40, seed=3):
    rng = np.random.default_rng(seed)
    t = np.sort(10.0 * rng.random(N))
    f = np.arange(Nf) * 0.5
    y = np.sin(2 * np.pi * 3.0 * t) + 0.6 * np.cos(2 * np.pi * 7.5 * t + 0.3) + 0.1 * rng.standard_normal(N)
    return t, f, y


@pytest.mark.parametrize("mode", ["gram", "literal"])
def test_l1_admm_limit_satisfies_lasso_kkt_and_matches_sklearn(mode):
    from sklearn.linear_model import Lasso

    t, f, y = _fourier_problem()
    lam = 0.5
    x, _, info = o.ls_sparse_spectral(y, t, f, lam=lam, iters=20000, tol=1e-11, mode=mode, return_info=True,
                                      printerval=10 ** 9)
    A, z = info["A"], info["z"]
    assert info["residual"] < 1e-11
    grad = A.T @ (y - A @ z)
    on = z != 0
    assert 0 < on.sum() < len(z)
    assert np.max(np.abs(grad[~on])) <= lam * (1 + 1e-7)          # |Aᵀr| ≤ λ off the support
    assert np.max(np.abs(grad[on] - lam * np.sign(z[on]))) <= 1e-7  # = λ sign(z) on it
    # an unrelated solver (coordinate descent) on the same problem: ½‖Ax−y‖²+λ‖x‖₁ = N·[1/(2N)‖·‖² + (λ/N)‖x‖₁]
    sk = Lasso(alpha=lam / len(y), fit_intercept=False, tol=1e-14, max_iter=200000).fit(A, y).coef_
    obj_admm = o.sparse_objective(A, y, z, o.NormL1(lam))
    obj_sk = o.sparse_objective(A, y, sk, o.NormL1(lam))
    assert abs(obj_admm - obj_sk) <= 1e-8 * max(1.0, abs(obj_sk))
    assert set(np.flatnonzero(z)) == set(np.flatnonzero(sk))
    assert np.linalg.norm(z - sk) <= 1e-6 * np.linalg.norm(sk)


def test_weighted_sparse_fits_minus_y_Q13():
    # Quadratic(Q,q) has the linear term +qᵀx (src/lasso.jl:120-121): with W = 1 the weighted method returns
    # exactly the negated coefficients of the unweighted one
    t, f, y = _fourier_problem(seed=4)
    a, _ = o.ls_sparse_spectral(y, t, f, lam=0.5, iters=3000, tol=1e-10, printerval=10 ** 9)
    b, _ = o.ls_sparse_spectral(y, t, f, np.ones(len(y)), lam=0.5, iters=3000, tol=1e-10, printerval=10 ** 9)
    assert np.count_nonzero(a) > 0 and np.linalg.norm(a + b) <= 1e-12 * np.linalg.norm(a)


def test_group_lasso_admm_limit_satisfies_kkt():
    y, v, x = o.generate_lpv_signal(400, seed=5)
    w = 2 * np.pi * np.arange(1, 13) * 2.0
    Nv, lam = 4, 6.0
    se, info = o.ls_sparse_spectral_lpv(y, x, v, w, Nv, lam=lam, iters=40000, tol=1e-11, return_info=True,
                                        printerval=10 ** 9)
    Phi, z = info["Phi"], info["z"]
    assert info["residual"] < 1e-11
    grad = (Phi.T @ (y - Phi @ z)).reshape(len(w), 2 * Nv)
    zg = z.reshape(len(w), 2 * Nv)
    nz = np.linalg.norm(zg, axis=1)
    assert 0 < np.count_nonzero(nz) < len(w)
    for g in range(len(w)):
        if nz[g] == 0:
            assert np.linalg.norm(grad[g]) <= lam * (1 + 1e-7)
        else:
            assert np.linalg.norm(grad[g] - lam * zg[g] / nz[g]) <= 1e-6
    # the un-permutation (src/lasso.jl:67) puts group f back at columns f + k·Nf
    P = se.x.reshape(len(w), Nv, order="F")
    assert np.array_equal(np.flatnonzero(np.abs(P).sum(axis=1)), np.flatnonzero(nz))


def test_l0_and_ball_admm_iterates_are_consistent():
    t, f, y = _fourier_problem(seed=6)
    for pg in (o.NormL0(0.5), o.IndBallL0(5)):
        _, _, info = o.ls_sparse_spectral(y, t, f, proxg=pg, iters=500, tol=1e-9, return_info=True,
                                          printerval=10 ** 9)
        z = info["z"]
        assert np.array_equal(pg.prox(z, 0.05), z)  # z is always a prox output: idempotent
        if isinstance(pg, o.IndBallL0):
            assert np.count_nonzero(z) <= 5


# ---------------------------------------------------------------------------------------------------------------
# basis: the regressor against a 50-digit evaluation of the phase the reference rounds (Q3)
# ---------------------------------------------------------------------------------------------------------------


def test_fourier_regressor_matches_mpmath_at_reference_rounding():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rng = np.random.default_rng(7)
    t = np.sort(10.0 * rng.random(12))
    f = np.array([0.0, 0.37, 11.0, 173.3, 2047.9])
    A, zf = o.get_fourier_regressor(t, f)
    assert zf == 0 and A.shape == (12, 9)
    dd = 1 / math.sqrt(2 * len(f))
    for k, fk in enumerate(f):
        for n, tn in enumerate(t):
            phi = (2 * math.pi * fk) * tn  # the reference's own FP64 rounding: fl(fl(2πf)·t)
            assert abs(A[n, k] - float(mp.cos(mp.mpf(phi))) * dd) <= 2e-16
            if k > 0:
                assert abs(A[n, len(f) - 1 + k] + float(mp.sin(mp.mpf(phi))) * dd) <= 2e-16


def test_hanning_is_dsp_symmetric_with_zero_ends():
    for n in (2, 5, 64, 4096):
        w = o.hanning(n)
        k = np.arange(n)
        assert w[0] == 0 and abs(w[-1]) <= 1e-16 and np.allclose(w, w[::-1], atol=1e-15)
        assert np.allclose(w, 0.5 * (1 - np.cos(2 * np.pi * k / (n - 1))), atol=1e-15)


# ---------------------------------------------------------------------------------------------------------------
# windows: brute-force restatement of DSP.arraysplit (k = (N−n)÷(n−noverlap)+1, hop n−noverlap, remainder dropped)
# ---------------------------------------------------------------------------------------------------------------


def test_window_offsets_bruteforce():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=300, deadline=None)
    @hyp.given(st.integers(1, 400), st.integers(1, 120), st.data())
    def run(N, n, data):
        nov = data.draw(st.integers(0, n - 1))
        starts = []
        s = 0
        while s + n <= N:
            starts.append(s)
            s += n - nov
        assert o.arraysplit_count(N, n, nov) == len(starts)
        assert list(o.window_offsets(N, n, nov)) == starts

    run()


def test_default_freqs_three_argument_form_uses_first_n_samples():
    # src/lsfft.jl:7-9: default_freqs(t, n) takes fs from t[1:n] only
    rng = np.random.default_rng(8)
    t = np.concatenate([np.sort(rng.random(64)), 1 + 10 * np.sort(rng.random(64))])
    f = o.default_freqs(t, 64)
    fs = 1 / np.mean(np.diff(t[:64]))
    assert len(f) == 33 and f[0] == 0 and abs(f[1] - fs / 64) <= 1e-12 * fs


# ---------------------------------------------------------------------------------------------------------------
# the oracle's two solve modes (what the reference runs / what the GPU runs) agree over random well-posed shapes
# ---------------------------------------------------------------------------------------------------------------


def test_literal_and_gram_modes_agree_over_random_shapes():
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=40, deadline=None, derandomize=True)
    @hyp.given(st.integers(0, 10 ** 6), st.integers(60, 260), st.integers(2, 14), st.booleans(), st.booleans())
    def run(seed, N, Nf, zero_first, weighted):
        rng = np.random.default_rng(seed)
        t = np.sort(10.0 * rng.random(N))
        f = (np.arange(Nf) + (0 if zero_first else 1)) * 0.37  # spacing 3.7 / record length: well conditioned
        y = rng.standard_normal(N)
        W = 0.2 + rng.random(N) if weighted else None
        a, _ = o.ls_spectral(y, t, f, W, lam=1e-10, mode="literal")
        b, _ = o.ls_spectral(y, t, f, W, lam=1e-10, mode="gram")
        assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(a)
        # windowed estimators: the ragged tail is dropped the same way in both modes, cohere(y,y) == 1 exactly (Q8)
        u = np.roll(y, 3) + 0.5 * rng.standard_normal(N)
        fw = f[:max(2, Nf // 2)] * 4
        for fn in (o.ls_windowcsd, o.ls_cohere):
            sa, _ = fn(y, u, t, fw, nw=3, mode="literal")
            sb, _ = fn(y, u, t, fw, nw=3, mode="gram")
            assert np.linalg.norm(sa - sb) <= 1e-8 * np.linalg.norm(sa)
        one, _ = o.ls_cohere(y, y, t, fw, nw=3, mode="gram")
        assert np.all(one == 1.0)

    run()


@pytest.mark.parametrize("N,Nf,f0,df", [(700, 33, 0.0, 0.31), (512, 20, 0.7, 0.45), (300, 1, 0.0, 1.0), (900, 65, 2.0, 0.05)])
def test_gram_from_trig_sums_is_the_product_form(N, Nf, f0, df):
    """The Toeplitz + Hankel identities behind the library's opt-in LPVS_PHASE_STRUCTURED mode (csrc/structured.cu), restated in
    the oracle: A'WA and A'Wy from 3 Nf trigonometric sums equal the products formed from get_fourier_regressor."""
    rng = np.random.default_rng(N + Nf)
    t = np.sort(10 * rng.random(N))
    y = rng.standard_normal(N)
    f = f0 + df * np.arange(Nf)
    A, _ = o.get_fourier_regressor(t, f)
    for W in (None, 0.5 + rng.random(N)):
        Aw = A if W is None else A * W[:, None]
        G, b = o.gram_from_trig_sums(t, f, W, y)
        assert G.shape == (A.shape[1], A.shape[1])
        assert np.abs(G - Aw.T @ A).max() <= 1e-13 * np.abs(Aw.T @ A).max()
        assert np.abs(b - Aw.T @ y).max() <= 1e-13 * np.abs(Aw.T @ y).max()


@pytest.mark.parametrize("half", [False, True])
def test_first_order_phase_correction_recovers_the_reference_gram(half):
    """LPVS_PHASE_STRUCTURED_REF (csrc/corr.cu) restated: at large phases (2.6e7 rad: the reference's fl(fl(2 pi f) t) is off the
    ideal phase by up to 2.9e-9 rad) the Gram matrix / right-hand side of the ideal grid differ from the reference's products by
    ~1e-10; adding the first-order terms D'B + B'D, D'y leaves <= 1e-3 of that -- with float16 operands too (what the tensor-core
    kernel uses), because only the DIFFERENCE goes through half precision."""
    rng = np.random.default_rng(12)
    N, Nf = 2048, 48
    t = 9.99 + np.sort(2.4e-3 * rng.random(N))
    f = np.arange(Nf) * (4096 / 2.4e-3 / 4096) * 5.0
    W = 3.0 * o.hanning(N)
    y = rng.standard_normal(N)
    A, _ = o.get_fourier_regressor(t, f)
    Gr, br = (A.T * W) @ A, (A.T * W) @ y
    G0, b0 = o.gram_from_trig_sums(t, f, W, y)
    dG, db = o.gram_phase_correction(t, f, W, y, half=half)
    e0 = np.abs(G0 - Gr).max() / np.abs(Gr).max()
    e1 = np.abs(G0 + dG - Gr).max() / np.abs(Gr).max()
    f0 = np.abs(b0 - br).max() / np.abs(br).max()
    f1 = np.abs(b0 + db - br).max() / np.abs(br).max()
    assert e0 > 1e-11 and f0 > 1e-11  # the effect is there
    assert e1 <= 1e-3 * e0 + 2e-14 and f1 <= 1e-3 * f0 + 2e-14, (e0, e1, f0, f1)


@pytest.mark.parametrize("f0,df,Nf,toff", [(3.1e3, 977.3, 33, 9.99), (0.0, 1234.5, 20, 55.0), (512.25, 2048.0, 65, 3.0)])
def test_first_order_phase_correction_on_other_grids(f0, df, Nf, toff):
    """The same statement for grids that do not start at zero (no dropped -sin column, Z+ != Z-) and other phase magnitudes:
    sums + first-order terms reproduce the reference-rounded products to <= 1e-3 of the uncorrected difference."""
    rng = np.random.default_rng(int(f0) + Nf)
    N = 1500
    t = toff + np.sort(3.0e-3 * rng.random(N))
    f = f0 + df * np.arange(Nf)
    W = 0.25 + rng.random(N)
    y = rng.standard_normal(N)
    A, _ = o.get_fourier_regressor(t, f)
    Gr, br = (A.T * W) @ A, (A.T * W) @ y
    G0, b0 = o.gram_from_trig_sums(t, f, W, y)
    dG, db = o.gram_phase_correction(t, f, W, y, half=True)
    e0 = np.abs(G0 - Gr).max() / np.abs(Gr).max()
    e1 = np.abs(G0 + dG - Gr).max() / np.abs(Gr).max()
    f0e = np.abs(b0 - br).max() / np.abs(br).max()
    f1e = np.abs(b0 + db - br).max() / np.abs(br).max()
    assert e1 <= 1e-3 * e0 + 3e-14 and f1e <= 1e-3 * f0e + 3e-14, (e0, e1, f0e, f1e)
