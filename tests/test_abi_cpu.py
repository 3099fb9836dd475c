"""The C-ABI library loads and exports every symbol include/lpvs.h declares; host-only entry points behave.
No compute calls (there is no GPU here).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import lpvs_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "lpvspectral.jl_b200", "liblpvs.so")
    if not os.path.exists(so):
        import __graft_entry__ as g

        g.build()
    from lpvspectral_jl_b200 import _lib

    return _lib


def test_exports_every_declared_symbol(lib):
    l = lib.load()
    syms = lib.declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/lpvs.h but not exported"
    assert set(syms) == set(lib._PROTOS), "ctypes prototypes out of sync with the header"
    assert l.lpvs_version() >= 100


def test_window_count_matches_arraysplit(lib):
    l = lib.load()
    for N, n, nov in [(100, 10, 0), (100, 10, 1), (1000, 125, 62), (5, 10, 0), (4194304, 4096, 2048), (10, 10, 9)]:
        assert l.lpvs_window_count(N, n, nov) == o.arraysplit_count(N, n, nov)
    assert l.lpvs_window_count(100, 10, -1) == o.arraysplit_count(100, 10, 5)
    assert l.lpvs_window_count(1 << 22, 4096, -1) == 2047
    assert l.lpvs_window_count(1 << 24, 4096, -1) == 8191


def test_finalize_matches_reference_normalisation(lib):
    import lpvspectral_jl_b200 as lp

    rng = np.random.default_rng(0)
    Nf, K = 17, 9
    s = rng.random(Nf)
    assert np.array_equal(lp.window_finalize(lib.WIN_PSD, s, Nf, K), s / float(K) ** 2)
    s2 = rng.standard_normal(2 * Nf)
    out = lp.window_finalize(lib.WIN_CSD, s2, Nf, K)
    assert np.array_equal(out, s2[:Nf] / K + 1j * (s2[Nf:] / K))
    syy, suu = rng.random(Nf) + 1, rng.random(Nf) + 1
    sr, si = rng.standard_normal(Nf), rng.standard_normal(Nf)
    out = lp.window_finalize(lib.WIN_COHERE, np.concatenate([syy, suu, sr, si]), Nf, K)
    assert np.array_equal(out, (sr * sr + si * si) / (suu * syy))
    # identical channels -> exactly 1
    out = lp.window_finalize(lib.WIN_COHERE, np.concatenate([syy, syy, syy, 0 * syy]), Nf, K)
    assert np.all(out == 1)


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import lpvspectral_jl_b200 as lp

    if lib.load().lpvs_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(lp.LpvsError):
        lp.Context(0)
    with pytest.raises(lp.LpvsError):
        lp.ls_spectral(np.ones(8), np.arange(8.0), np.array([0.0, 0.1]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lpvspectral.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as fh:
                    src = fh.read()
                assert "oracle" not in src.replace("# oracle", ""), f"{fn} references the oracle"


def test_host_helpers_match_oracle():
    import lpvspectral_jl_b200 as lp

    t = np.sort(10 * np.random.default_rng(3).random(500))
    assert np.array_equal(lp.default_freqs(t), o.default_freqs(t))
    assert np.array_equal(lp.default_freqs(t, 100), o.default_freqs(t, 100))
    for n in (1, 2, 7, 64):
        assert np.array_equal(lp.hanning(n), o.hanning(n))
        assert np.array_equal(lp.hamming(n), o.hamming(n))
        assert np.array_equal(lp.rect(n), o.rect(n))
    assert lp.check_freq(np.array([0.0, 1.0])) == 0 and lp.check_freq(np.array([1.0, 2.0])) is None
    with pytest.raises(ValueError):
        lp.check_freq(np.array([1.0, 0.0]))


def test_plain_c_caller_compiles_links_and_fails_loudly_without_gpu(lib, tmp_path):
    """include/lpvs.h is valid C99 and liblpvs.so links from a plain C program (examples/c_abi_example.c): host-only
    entry points work, and without a GPU lpvs_init refuses (exit code 3) instead of computing on the host."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pkg = os.path.join(ROOT, "lpvspectral.jl_b200")
    exe = str(tmp_path / "lpvs_example")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                    os.path.join(ROOT, "include", "lpvs.h")], check=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_example.c"), "-L" + pkg, "-llpvs", "-lm",
                    "-Wl,-rpath," + pkg, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert "15 windows at 50% overlap" in out.stdout
    if lib.load().lpvs_device_count() > 0:
        assert out.returncode == 0, out.stdout + out.stderr
        assert "peak at f = 20.00 Hz" in out.stdout
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stdout
