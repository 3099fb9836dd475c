"""Pins the oracle on every deterministic known-answer test the reference holds for this path
(/root/reference/test/runtests.jl:27-58,168-208).  CPU only."""
import numpy as np
import pytest

from oracle import lpvs_oracle as o


def test_windows_counts_and_offsets():
    # test/runtests.jl:27-58
    y = np.arange(1, 101)
    W = o.Windows((y, y), 10, 0)
    assert len(W) == 10
    first = next(iter(W))
    assert np.array_equal(first[0], np.arange(1, 11)) and np.array_equal(first[1], np.arange(1, 11))
    W = o.Windows((y, y), 10, 1)
    assert len(W) == 11
    cW = list(W)
    assert np.array_equal(cW[0][0], np.arange(1, 11))
    assert np.array_equal(cW[1][0], np.arange(10, 20))
    W3 = o.Windows((y, y, y), 10, 1)
    assert len(W3) == 11 and np.array_equal(list(W3)[1][2], np.arange(10, 20))


@pytest.mark.parametrize("nov", [0, 1])
def test_mapwindows_roundtrip(nov):
    # test/runtests.jl:32-35,45-48: mapwindows(-y) == -1:-1:-100
    y = np.arange(1, 101).astype(float)
    W = o.Windows((y, y), 10, nov)
    res = o.merge_windows([-w[0] for w in W], 100, 10, nov)
    assert np.array_equal(res, -y)


def test_default_freqs_and_check_freq():
    # test/runtests.jl:168-183
    t = np.arange(1000) * 0.1
    f = o.default_freqs(t)
    assert f[0] == 0 and abs(f[-1] - 5) < 1e-12 and len(f) == 501
    assert o.check_freq(f) == 0
    with pytest.raises(ValueError):
        o.check_freq(np.array([1.0, 0.0, 2.0]))
    A, z = o.get_fourier_regressor(t, f)
    assert A.shape == (1000, 2 * len(f) - 1)


def test_ls_kats():
    # test/runtests.jl:186-208
    t = np.arange(1000) * 0.1
    f = o.default_freqs(t)
    y = np.sin(2 * np.pi * t)
    x, fr = o.ls_spectral(y, t)
    a = o._abs2(x)
    assert a.argmax() + 1 == 101 and abs(a.max() - 2.0 * len(fr)) < 1e-4
    x, _ = o.ls_spectral(y, t, f, np.ones(len(y)))
    a = o._abs2(x)
    assert a.argmax() + 1 == 101 and abs(a.max() - 2.0 * len(fr)) < 1e-4
    S, fr = o.ls_windowpsd(y, t, noverlap=0)
    assert S.argmax() + 1 == 13
    S, fr = o.ls_windowpsd(y, t, nw=16, noverlap=0)
    assert np.abs(S).argmax() + 1 == 7
    S, fr = o.ls_windowcsd(y, y, t, noverlap=0)
    assert np.abs(S).argmax() + 1 == 11 and abs(np.abs(S).max() - 2.0 * len(fr)) < 1e-4
    Cxy, _ = o.ls_cohere(y, y, t)
    assert np.all(Cxy == 1)


def test_cohere_noisy_structure():
    # test/runtests.jl:210-213 (RNG-dependent in the reference; structure reused with a numpy seed)
    rng = np.random.default_rng(0)
    t = np.arange(1000) * 0.1
    y = np.sin(2 * np.pi * t)
    Cxy, _ = o.ls_cohere(y, y + 0.5 * rng.standard_normal(len(y)), t, nw=8, noverlap=-1)
    assert abs(Cxy.max() - 1.0) < 0.15 and abs(int(Cxy.argmax()) + 1 - 14) <= 1
    assert Cxy.mean() < 0.25


def test_literal_and_gram_modes_agree():
    rng = np.random.default_rng(1)
    t = np.sort(10 * rng.random(1024))
    y = np.sin(2 * np.pi * 20 * t) + 0.1 * rng.standard_normal(1024)
    f = o.default_freqs(t)[:128]
    a, _ = o.ls_spectral(y, t, f, mode="literal")
    b, _ = o.ls_spectral(y, t, f, mode="gram")
    assert np.linalg.norm(a - b) <= 1e-11 * np.linalg.norm(a)
    W = o.hanning(1024)
    a, _ = o.ls_spectral(y, t, f[::2], W, mode="literal")
    b, _ = o.ls_spectral(y, t, f[::2], W, mode="gram")
    assert np.linalg.norm(a - b) <= 1e-11 * np.linalg.norm(a)


def test_prox_operators():
    v = np.array([3.0, -0.2, 0.5, -4.0, 0.0, 1.0])
    assert np.allclose(o.NormL1(2.0).prox(v, 0.5), [2.0, 0.0, 0.0, -3.0, 0.0, 0.0])
    assert np.array_equal(o.NormL0(0.5).prox(v, 1.0), [3.0, 0.0, 0.0, -4.0, 0.0, 0.0])  # |v| > 1
    assert np.array_equal(o.IndBallL0(2).prox(v, 1.0), [3.0, 0.0, 0.0, -4.0, 0.0, 0.0])
    g = o.GroupNormL2(1.0, 3, 2).prox(v, 1.0)
    n1, n2 = np.linalg.norm(v[:3]), np.linalg.norm(v[3:])
    assert np.allclose(g[:3], (1 - 1 / n1) * v[:3]) and np.allclose(g[3:], (1 - 1 / n2) * v[3:])
    # ties in IndBallL0 -> lower index
    assert np.array_equal(o.IndBallL0(1).prox(np.array([1.0, -1.0]), 1.0), [1.0, 0.0])


def test_lpv_group_perm_matches_reference_formula():
    # src/lasso.jl:47: inds = reshape(1:len, Nf, :)'[:]
    Nf, Nv = 3, 2
    n = 2 * Nf * Nv
    inds = o.lpv_group_perm(Nf, n)
    expect = np.arange(1, n + 1).reshape(-1, Nf).T.ravel() - 1  # reshape(1:n, Nf, :) is column-major
    assert np.array_equal(inds, expect)
    for f in range(Nf):
        grp = inds[f * 2 * Nv:(f + 1) * 2 * Nv]
        assert np.all(grp % Nf == f)
