"""Direct ctypes calls into entry points the Python mirror does not route through: the one-shot sparse functions,
a caller-supplied ADMM start vector, and the lpvs_dev_* helpers a CUDA-less host (the Julia shim) would use to keep
inputs resident.  (-m gpu)"""
import ctypes as C

import numpy as np
import pytest

from oracle import lpvs_oracle as o

pytestmark = pytest.mark.gpu


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def sig(N, seed):
    rng = np.random.default_rng(seed)
    t = np.sort(10 * rng.random(N))
    y = np.sin(2 * np.pi * 3 * t) + 0.6 * np.cos(2 * np.pi * 8.5 * t) + 0.1 * rng.standard_normal(N)
    return t, y


def test_one_shot_sparse_functions(ctx):
    from lpvspectral_jl_b200 import _lib as L

    t, y = sig(700, 1)
    f = np.arange(0, 120) * 0.1
    x = np.empty(len(f), dtype=np.complex128)
    its, res = C.c_int64(0), C.c_double(0)
    ctx.check(ctx.lib.lpvs_ls_sparse_spectral(ctx.h, vp(y), vp(t), len(y), vp(f), len(f), None, L.PROX_L1, 0.3, 0.05,
                                              0, 0.0, 2000, 1e-9, vp(x), C.byref(its), C.byref(res)))
    xr, _, ri = o.ls_sparse_spectral(y, t, f, lam=0.3, iters=2000, tol=1e-9, mode="gram", return_info=True,
                                     printerval=10 ** 9)
    assert its.value == ri["iters"] and rel(x, xr) <= 1e-9 and res.value < 1e-9

    Y, V, X = o.generate_lpv_signal(300, seed=7)
    w = 2 * np.pi * np.arange(2, 12, 2)
    p = np.empty(len(w) * 6, dtype=np.complex128)
    ctx.check(ctx.lib.lpvs_ls_sparse_spectral_lpv(ctx.h, vp(Y), vp(X), vp(V), len(Y), vp(w), len(w), 6, 0, 1, 1.5, 0.05,
                                                  800, 1e-8, vp(p), C.byref(its), C.byref(res)))
    sr, ri = o.ls_sparse_spectral_lpv(Y, X, V, w, 6, lam=1.5, iters=800, tol=1e-8, mode="gram", return_info=True,
                                      printerval=10 ** 9)
    assert its.value == ri["iters"] and rel(p, sr.x) <= 1e-9


def test_admm_with_caller_start_vector(ctx):
    from lpvspectral_jl_b200 import _lib as L
    import lpvspectral_jl_b200 as lp

    t, y = sig(500, 2)
    f = np.arange(0, 80) * 0.15  # zero frequency first -> Nreg = 159
    nreg = 2 * len(f) - 1
    x0 = 0.1 * np.random.default_rng(0).standard_normal(nreg)
    h = C.c_void_p()
    ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, vp(y), vp(t), len(y), vp(f), len(f), None, L.PROX_L1, 0.2, 0.05,
                                               vp(x0), 0, 0.0, C.byref(h)))
    s = lp.ADMM(ctx, h)
    assert s.size == nreg
    xg, zg = s.get()
    assert np.array_equal(xg, x0) and np.array_equal(zg, x0)  # z = copy(x) (src/lasso.jl:146)
    s.step(300, 0.0)
    xg, zg = s.get()
    s.free()
    A, _ = o.get_fourier_regressor(t, f)
    pf = o.QuadProx(A.T @ A, A.T @ y, "ls", "gram")
    xr, zr, _, _ = o.admm(x0, pf, o.NormL1(0.2), iters=300, tol=0.0, mu=0.05, printerval=10 ** 9)
    assert rel(zg, zr) <= 1e-9 and rel(xg, xr) <= 1e-9


def test_resident_inputs_via_lpvs_dev_helpers(ctx):
    """What a host without CUDA bindings does: lpvs_dev_alloc/upload once, then many *_dev calls."""
    from lpvspectral_jl_b200 import _lib as L
    import lpvspectral_jl_b200 as lp

    t, y = sig(8192, 3)
    n, nf = 1024, 32
    f = np.arange(nf) * 2.0 / (t[n] - t[0])
    W = lp.hanning(n)
    dy, dt = C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.lpvs_dev_alloc(ctx.h, y.nbytes, C.byref(dy)))
    ctx.check(ctx.lib.lpvs_dev_alloc(ctx.h, t.nbytes, C.byref(dt)))
    ctx.check(ctx.lib.lpvs_dev_upload(ctx.h, dy, vp(y), y.nbytes))
    ctx.check(ctx.lib.lpvs_dev_upload(ctx.h, dt, vp(t), t.nbytes))
    back = np.empty_like(y)
    ctx.check(ctx.lib.lpvs_dev_download(ctx.h, vp(back), dy, y.nbytes))
    assert np.array_equal(back, y)
    K = lp.window_count(len(y), n, -1)
    sums = np.zeros(nf)
    info = C.c_int(0)
    ctx.check(ctx.lib.lpvs_ls_window_sums_dev(ctx.h, L.WIN_PSD, dy, None, dt, len(y), vp(f), nf, vp(W), n, n >> 1,
                                              1e-10, 0, K, vp(sums), C.byref(info)))
    S = lp.window_finalize(L.WIN_PSD, sums, nf, K)
    Sr, _ = lp.ls_windowpsd(y, t, f, nw=len(y) // n, window_func=lp.hanning, ctx=ctx)
    assert np.array_equal(S, Sr)
    ms = ctx.last_call_ms()
    assert ms > 0 and ctx.launches > 0
    ctx.check(ctx.lib.lpvs_dev_free(ctx.h, dy))
    ctx.check(ctx.lib.lpvs_dev_free(ctx.h, dt))


def test_bad_arguments_return_codes(ctx):
    from lpvspectral_jl_b200 import _lib as L

    t, y = sig(64, 4)
    f = np.array([1.0, 2.0])
    x = np.empty(2, dtype=np.complex128)
    info = C.c_int(0)
    assert ctx.lib.lpvs_ls_spectral(ctx.h, None, vp(t), 64, vp(f), 2, None, 1e-10, vp(x), C.byref(info)) == L.E_BAD_ARG
    assert ctx.lib.lpvs_ls_spectral(None, vp(y), vp(t), 64, vp(f), 2, None, 1e-10, vp(x), C.byref(info)) == L.E_BAD_ARG
    sums = np.zeros(2)
    W = np.ones(16)
    rc = ctx.lib.lpvs_ls_window_sums(ctx.h, 7, vp(y), None, vp(t), 64, vp(f), 2, vp(W), 16, 8, 1e-10, 0, 1, vp(sums),
                                     C.byref(info))
    assert rc == L.E_BAD_ARG
    rc = ctx.lib.lpvs_ls_window_sums(ctx.h, L.WIN_PSD, vp(y), None, vp(t), 64, vp(f), 2, vp(W), 16, 8, 1e-10, 0, 99,
                                     vp(sums), C.byref(info))
    assert rc == L.E_BAD_ARG and b"window range" in ctx.lib.lpvs_last_error(ctx.h)
    h = C.c_void_p()
    rc = ctx.lib.lpvs_admm_create_fourier(ctx.h, vp(y), vp(t), 64, vp(f), 2, None, L.PROX_L1, 0.1, 0.0, None, 0, 0.0,
                                          C.byref(h))
    assert rc == L.E_BAD_ARG  # mu must be in (0, 1]
    # windowed sparse entry point: same validation sites
    sp = lambda kind, u, prox, mu, k1: ctx.lib.lpvs_ls_window_sparse_sums(  # noqa: E731
        ctx.h, kind, vp(y), u, vp(t), 64, vp(f), 2, vp(W), 16, 8, prox, 0.1, mu, 10, 1e-6, 0, k1, vp(sums), None, None,
        C.byref(info))
    assert sp(L.WIN_PSD, None, L.PROX_GROUP_L2, 0.05, 1) == L.E_BAD_ARG      # group prox is LPV only
    assert sp(L.WIN_PSD, None, L.PROX_L1, 1.5, 1) == L.E_BAD_ARG             # mu in (0, 1]
    assert sp(L.WIN_CSD, None, L.PROX_L1, 0.05, 1) == L.E_BAD_ARG            # second signal required
    assert sp(L.WIN_PSD, None, L.PROX_L1, 0.05, 99) == L.E_BAD_ARG           # window range
    assert sp(L.WIN_PSD, None, L.PROX_L1, 0.05, 1) == L.OK


def test_windowed_sparse_direct_call_outputs(ctx):
    """lpvs_ls_window_sparse_sums through ctypes: per-window iteration counts / residuals come back per channel, a
    zero-iteration budget leaves z = x0 = 0 (sums exactly zero), and tol = inf stops every window after one iteration."""
    from lpvspectral_jl_b200 import _lib as L

    t, y = sig(1200, 9)
    u = np.roll(y, 3).copy()
    f = np.arange(1, 20) * 0.6
    n, nov = 200, 100
    W = np.hanning(n)
    K = int(ctx.lib.lpvs_window_count(len(y), n, nov))
    sums = np.ones(4 * len(f))
    its = np.full(2 * K, -1, dtype=np.int64)
    res = np.full(2 * K, -1.0)
    info = C.c_int(0)
    call = lambda iters, tol: ctx.lib.lpvs_ls_window_sparse_sums(  # noqa: E731
        ctx.h, L.WIN_COHERE, vp(y), vp(u), vp(t), len(y), vp(f), len(f), vp(W), n, nov, L.PROX_L1, 0.05, 0.05, iters,
        tol, 0, K, vp(sums), vp(its), vp(res), C.byref(info))
    assert call(0, 1e-9) == L.OK
    assert np.all(sums == 0.0) and np.all(its == 0)
    assert call(50, np.inf) == L.OK
    assert np.all(its == 1) and np.all(res >= 0.0) and np.any(sums != 0.0)
    assert call(300, 1e-7) == L.OK
    assert its.min() >= 2 and its.max() <= 300 and np.all(res[its < 300] < 1e-7)


def test_sharded_admm_entry_points_single_gpu(ctx):
    """lpvs_admm_shard_*: argument validation, handle export, and the guard against running unconnected.  The exchange
    itself needs >= 2 GPUs (tests/test_gpu_multi.py, tools/admm_shard_check.py)."""
    from lpvspectral_jl_b200 import _lib as L

    t, y = sig(900, 6)
    f = np.arange(0, 300) * 0.1

    def create(prox=L.PROX_L1, param=0.2):
        h = C.c_void_p()
        ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, vp(y), vp(t), len(y), vp(f), len(f), None, prox, param, 0.05,
                                                   None, 0, 0.0, C.byref(h)))
        return h

    h = create()
    assert ctx.lib.lpvs_admm_shard_begin(h, 0, 1) == L.E_BAD_ARG        # world >= 2
    assert ctx.lib.lpvs_admm_shard_begin(h, 2, 2) == L.E_BAD_ARG        # rank < world
    assert ctx.lib.lpvs_admm_shard_begin(h, 0, 8) == L.E_UNSUPPORTED    # 5 row blocks cannot feed 8 ranks
    buf = C.create_string_buffer(64)
    assert ctx.lib.lpvs_admm_shard_handle(h, C.cast(buf, C.c_void_p)) == L.E_BAD_ARG  # not sharded yet
    assert ctx.lib.lpvs_admm_shard_begin(h, 1, 2) == L.OK
    assert ctx.lib.lpvs_admm_shard_handle(h, C.cast(buf, C.c_void_p)) == L.OK and any(buf.raw)
    it, res, conv = C.c_int64(0), C.c_double(0.0), C.c_int(0)
    rc = ctx.lib.lpvs_admm_run(h, 10, 1e-9, C.byref(it), C.byref(res), C.byref(conv))
    assert rc == L.E_BAD_ARG and b"not connected" in ctx.lib.lpvs_last_error(ctx.h)
    assert ctx.lib.lpvs_admm_shard_begin(h, 1, 2) == L.E_UNSUPPORTED    # already sharded
    ctx.lib.lpvs_admm_free(h)
    # IndBallL0 and the group prox shard only with the one-exchange scheme (every rank then holds every row)
    h = create(L.PROX_BALL_L0, 5.0)
    ctx.set_option(L.OPT_SHARD_EXCHANGE, 0)
    try:
        assert ctx.lib.lpvs_admm_shard_begin(h, 0, 2) == L.E_UNSUPPORTED
    finally:
        ctx.set_option(L.OPT_SHARD_EXCHANGE, 2)
    assert ctx.lib.lpvs_admm_shard_begin(h, 0, 2) == L.OK
    ctx.lib.lpvs_admm_free(h)


def test_release_workspace_then_reuse(ctx):
    """lpvs_release_workspace frees the grow-only tables; the next call re-allocates and returns the same answer."""
    import lpvspectral_jl_b200 as lp

    t, y = sig(600, 11)
    f = np.arange(0, 40.0)
    x0 = lp.ls_spectral(y, t, f, ctx=ctx)[0]
    ctx.release_workspace()
    x1 = lp.ls_spectral(y, t, f, ctx=ctx)[0]
    assert np.array_equal(x0, x1)
    assert ctx.lib.lpvs_release_workspace(None) != 0
