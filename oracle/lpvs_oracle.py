"""CPU oracle for the LPVSpectral.jl least-squares spectral estimation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lpvspectral.jl_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and there only
as the checker or the CPU arm that is timed beside the GPU -- never as the product path.

It is a numpy/scipy FP64 *restatement* of the reference's algorithms (file:line citations are into
``/root/reference``).  Julia is not installed in the build image nor on the GPU box, so the reference itself
cannot be executed; the restatement is pinned instead by every deterministic known-answer test the reference
holds for this path (``test/runtests.jl:27-58,168-208`` -- see ``tests/test_oracle_kats.py``).

Third-party semantics that are NOT under ``/root/reference`` and are restated from their published
behaviour (unpinned by the reference's own tests, see DESIGN.md "parity pins"):

* ``DSP.arraysplit``/``hanning``/``rect`` (DSP.jl compat 0.7/0.8), ``FFTW.rfftfreq`` (FFTW.jl 1.1),
* LinearAlgebra: ``svd(M)\\b`` (singular values below ``eps*s[0]`` dropped), rectangular ``\\`` (pivoted QR),
  square ``\\`` (LU), ``var`` (n-1),
* ProximalOperators.jl (compat 0.10/0.15/0.16): ``LeastSquares(A,b,iterative=true)``, ``Quadratic(Q,q,
  iterative=true)``, ``NormL1``, ``NormL0``, ``IndBallL0``, ``NormL2``, ``SlicedSeparableSum``,
* IterativeSolvers.jl ``cg!`` (warm start, reltol=sqrt(eps) w.r.t. the initial residual, maxiter=n).

ADMM parity is therefore "unpinned" by the reference (``test/test_lasso.jl`` has no assertions): this file is
the de-facto specification for the prox steps.  What can be checked without Julia is checked in
``tests/test_oracle_pins.py``: every prox operator against a brute-force minimisation of its Moreau objective, the ADMM
limits against the lasso / group-lasso optimality conditions and scikit-learn, the CG restatement against scipy.

Two solve modes everywhere a dense solve happens:

* ``mode="literal"``: the algorithm the reference runs (SVD / N-RHS LU / pivoted QR / warm-started CG);
* ``mode="gram"``: Gram + Cholesky / exact x-update -- the algorithm the CUDA library runs, on the CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

EPS = np.finfo(np.float64).eps

# --------------------------------------------------------------------------------------------------------
# windows (DSP.jl semantics, restated) -- src/windows.jl:7-42,86-110
# --------------------------------------------------------------------------------------------------------


def rect(n: int) -> np.ndarray:
    """DSP.rect: ones(n)."""
    return np.ones(n)


def hanning(n: int) -> np.ndarray:
    """DSP.hanning(n): symmetric 0.5(1-cos(2 pi k/(n-1))), zero end weights."""
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * k / (n - 1)))


def hamming(n: int) -> np.ndarray:
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))


def arraysplit_count(N: int, n: int, noverlap: int) -> int:
    """DSP.arraysplit length: N >= n ? (N-n) div (n-noverlap) + 1 : 0 (pinned by test/runtests.jl:29-58)."""
    return (N - n) // (n - noverlap) + 1 if N >= n else 0


def window_offsets(N: int, n: int, noverlap: int) -> np.ndarray:
    """0-based start index of every window; window i covers [off_i, off_i+n)."""
    K = arraysplit_count(N, n, noverlap)
    return np.arange(K, dtype=np.int64) * (n - noverlap)


@dataclass
class Windows:
    """Windows2 / Windows3 (src/windows.jl:27-36, 94-104): ``noverlap < 0`` means ``n >> 1``."""

    signals: Tuple[np.ndarray, ...]
    n: int
    noverlap: int
    W: np.ndarray

    def __init__(self, signals: Sequence[np.ndarray], n: int, noverlap: int = -1, window_func: Callable = rect):
        sigs = tuple(np.asarray(s) for s in signals)
        N = len(sigs[0])
        for s in sigs:
            if len(s) != N:
                raise AssertionError("y and t has to be the same length")  # src/windows.jl:31,96
        if noverlap < 0:
            noverlap = n >> 1
        self.signals = sigs
        self.n = n
        self.noverlap = noverlap
        self.W = np.asarray(window_func(n), dtype=np.float64)
        self.offsets = window_offsets(N, n, noverlap)

    def __len__(self):
        return len(self.offsets)

    def __iter__(self):
        for o in self.offsets:
            yield tuple(s[o:o + self.n] for s in self.signals)


def merge_windows(pieces: Sequence[np.ndarray], N: int, n: int, noverlap: int) -> np.ndarray:
    """Base.merge(yf, w::Windows2) (src/windows.jl:58-70): overlap-average re-assembly used by mapwindows."""
    ym = np.zeros(N, dtype=np.asarray(pieces[0]).dtype)
    counts = np.zeros(N, dtype=np.int64)
    lo, hi = 0, n
    for p in pieces:
        hi_c = min(hi, N)
        ym[lo:hi_c] += np.asarray(p)[: hi_c - lo]
        counts[lo:hi_c] += 1
        lo += n - noverlap
        hi += n - noverlap
    return ym / np.maximum(counts, 1)


# --------------------------------------------------------------------------------------------------------
# frequency grid / Fourier regressor -- src/lsfft.jl:3-49
# --------------------------------------------------------------------------------------------------------


def default_freqs_n(n: int, fs: float = 1.0) -> np.ndarray:
    """default_freqs(n::Int, fs) (src/lsfft.jl:3-6): rfftfreq(n,fs) = (0:n div 2)*fs/n."""
    return np.arange(n // 2 + 1, dtype=np.float64) * (fs / n)


def default_freqs(t: np.ndarray, n: Optional[int] = None) -> np.ndarray:
    """default_freqs(t) / default_freqs(t, n::Int) (src/lsfft.jl:7,9): the 2-arg form uses t[1:n] only."""
    t = np.asarray(t, dtype=np.float64)
    if n is not None:
        t = t[:n]
    fs = 1.0 / np.mean(np.diff(t))
    return default_freqs_n(len(t), fs)


def check_freq(f: np.ndarray) -> Optional[int]:
    """src/lsfft.jl:20-24.  Returns the 0-based index of the zero frequency (only 0 allowed) or None."""
    f = np.asarray(f)
    z = np.flatnonzero(f == 0)
    if len(z) and z[0] != 0:
        raise ValueError("If zero frequency is included it must be the first frequency")
    return 0 if len(z) else None


def nreg_of(f: np.ndarray) -> int:
    return 2 * len(f) - (1 if check_freq(f) is not None else 0)


def get_fourier_regressor(t: np.ndarray, f: np.ndarray) -> Tuple[np.ndarray, Optional[int]]:
    """src/lsfft.jl:26-49: A[:,k]=cos(phi)*dd, A[:,k+off]=-sin(phi)*dd, phi=(2pi*f_k)*t_n, dd=1/sqrt(2Nf)."""
    t = np.asarray(t, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    zerofreq = check_freq(f)
    N, Nf = len(t), len(f)
    dd = 1.0 / math.sqrt(2 * Nf)
    phi = np.outer(t, (2.0 * np.pi) * f)  # fl(fl(2pi f) t), the reference's rounding order (Q3)
    C = np.cos(phi) * dd
    S = -np.sin(phi) * dd
    if zerofreq is not None:
        S = S[:, 1:]
    return np.concatenate([C, S], axis=1), zerofreq


def gram_from_trig_sums(t: np.ndarray, f: np.ndarray, W: Optional[np.ndarray] = None,
                        y: Optional[np.ndarray] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """A'WA and A'Wy of get_fourier_regressor on a UNIFORM grid f_k = f0 + k df from the 3 Nf trigonometric sums
    Z-(m) = sum_s W_s e^{-i 2 pi m df t_s}, Z+(m) = sum_s W_s e^{-i 2 pi (2 f0 + m df) t_s} and Zy(k) = sum_s W_s y_s e^{-i a_k(s)}
    (product-to-sum identities: the matrix is Toeplitz + Hankel in the frequency index).  CPU restatement of the opt-in
    LPVS_PHASE_STRUCTURED mode of the library (csrc/structured.cu) -- not a reference algorithm: src/lsfft.jl:77 forms the
    product A'*Wd*A.  Phases are those of the ideal grid f0 + k df (long double), not the reference's fl(fl(2 pi f) t)."""
    t = np.asarray(t, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    zerofreq = check_freq(f)
    N, Nf = len(t), len(f)
    Wv = np.ones(N) if W is None else np.asarray(W, dtype=np.float64)
    LD = np.longdouble
    f0 = LD(f[0])
    df = LD((f[-1] - f[0]) / (Nf - 1)) if Nf > 1 else LD(0.0)
    tl = t.astype(LD)

    def zsum(freqs, weights):  # sum_s w_s (cos, sin)(2 pi F t_s), the phase reduced in turns in long double
        turns = np.outer(freqs, tl)
        r = (turns - np.rint(turns)).astype(np.float64)
        return np.cos(2.0 * np.pi * r) @ weights, np.sin(2.0 * np.pi * r) @ weights

    m_minus = np.arange(Nf).astype(LD)
    m_plus = np.arange(2 * Nf - 1).astype(LD)
    Cm, Sm = zsum(m_minus * df, Wv)
    Cp, Sp = zsum(2 * f0 + m_plus * df, Wv)
    i = np.arange(Nf)[:, None]
    j = np.arange(Nf)[None, :]
    d, sgn = np.abs(i - j), np.sign(i - j)
    cc = 0.5 * (Cm[d] + Cp[i + j])
    ss = 0.5 * (Cm[d] - Cp[i + j])
    sc = -0.5 * (Sp[i + j] + sgn * Sm[d])  # rows -sin_i, columns cos_j
    cs = -0.5 * (Sp[i + j] - sgn * Sm[d])  # rows cos_i, columns -sin_j
    G = np.block([[cc, cs], [sc, ss]]) / (2 * Nf)
    b = None
    if y is not None:
        Cy, Sy = zsum(f0 + m_minus * df, Wv * np.asarray(y, dtype=np.float64))
        b = np.concatenate([Cy, -Sy]) / math.sqrt(2 * Nf)
    if zerofreq is not None:  # the -sin column of the zero frequency does not exist
        keep = np.r_[0:Nf, Nf + 1:2 * Nf]
        G = G[np.ix_(keep, keep)]
        b = None if b is None else b[keep]
    return G, b


def _two_prod_err(a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """p = fl(a b) and e = a b - p exactly (Dekker / Veltkamp splitting: what one FMA gives on the device)."""
    p = a * b
    c = 134217729.0  # 2^27 + 1
    ah = (a * c) - ((a * c) - a)
    al = a - ah
    bh = (b * c) - ((b * c) - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def gram_phase_correction(t: np.ndarray, f: np.ndarray, W: Optional[np.ndarray] = None, y: Optional[np.ndarray] = None,
                          half: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """First-order correction that takes gram_from_trig_sums (ideal phases theta = 2 pi (f0 + k df) t) to the reference's basis
    at phi = fl(fl(2 pi f) t) (src/lsfft.jl:34,41): with eps = phi - theta, B = (cos, -sin)(phi) and D = W eps (-sin, -cos)(phi),
        dG = (D'B + B'D) / (2 Nf),   db = D'y / sqrt(2 Nf).
    CPU restatement of LPVS_PHASE_STRUCTURED_REF (csrc/corr.cu) -- not a reference algorithm.  eps is exact: the rounding error
    of the product w t (two-product) plus (fl(2 pi f_k) - 2 pi f_k) t; `half` rounds eps and the operands to float16 as the
    tensor-core kernel does (float32 accumulation)."""
    t = np.asarray(t, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    zerofreq = check_freq(f)
    N, Nf = len(t), len(f)
    Wv = np.ones(N) if W is None else np.asarray(W, dtype=np.float64)
    LD = np.longdouble
    twopi = LD("6.283185307179586476925286766559005768")
    df = LD((f[-1] - f[0]) / (Nf - 1)) if Nf > 1 else LD(0.0)
    w = 2.0 * np.pi * f  # fl(2 pi f), src/lsfft.jl:34
    dw = (w.astype(LD) - twopi * (LD(f[0]) + df * np.arange(Nf).astype(LD))).astype(np.float64)
    p, e = _two_prod_err(t[:, None], w[None, :])  # p = fl(w t), e = w t - p
    eps = dw[None, :] * t[:, None] - e
    c, ms = np.cos(p), -np.sin(p)
    if half:
        scale = 2.0 ** (14 - math.frexp(max(np.abs(eps).max(), 1e-300))[1])
        wsc = 2.0 ** (-math.frexp(np.abs(Wv).max())[1])
        ef = (eps * scale).astype(np.float16).astype(np.float32)
        de = ef * (Wv * wsc).astype(np.float32)[:, None]
        q = lambda a: a.astype(np.float16).astype(np.float32)  # noqa: E731
        D = np.hstack([q(de * ms.astype(np.float32)), q(-de * c.astype(np.float32))])
        B = np.hstack([q(c), q(ms)])
        M = (D.T @ B).astype(np.float64) / (scale * wsc)
        db = None if y is None else (np.hstack([de * ms.astype(np.float32), -de * c.astype(np.float32)]).T.astype(np.float64)
                                     @ np.asarray(y, dtype=np.float64)) / (scale * wsc)
    else:
        de = eps * Wv[:, None]
        D = np.hstack([de * ms, -de * c])
        B = np.hstack([c, ms])
        M = D.T @ B
        db = None if y is None else D.T @ np.asarray(y, dtype=np.float64)
    dG = (M + M.T) / (2 * Nf)
    if db is not None:
        db = db / math.sqrt(2 * Nf)
    if zerofreq is not None:
        keep = np.r_[0:Nf, Nf + 1:2 * Nf]
        dG = dG[np.ix_(keep, keep)]
        db = None if db is None else db[keep]
    return dG, db


def fourier2complex(x: np.ndarray, zerofreq: Optional[int]) -> np.ndarray:
    """src/utilities.jl:62-73."""
    x = np.asarray(x)
    n = len(x) // 2
    if zerofreq is None:
        return x[:n] + 1j * x[n:]
    x0 = x[0]
    rest = x[1:]
    c = rest[:n] + 1j * rest[n:]
    return np.concatenate([[x0 + 0j], c])


def complex2fourier(params: np.ndarray, zerofreq: Optional[int]) -> np.ndarray:
    """The inverse packing used for ADMM start vectors (src/lasso.jl:93-97)."""
    if zerofreq is None:
        return np.concatenate([params.real, params.imag])
    return np.concatenate([params.real, params.imag[1:]])


# --------------------------------------------------------------------------------------------------------
# dense solves
# --------------------------------------------------------------------------------------------------------


def svd_solve(M: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Julia ``svd(M) \\ b``: drop singular values <= eps*s[0] (LinearAlgebra stdlib, restated)."""
    U, s, Vt = sla.svd(M, full_matrices=False, lapack_driver="gesdd")
    k = int(np.sum(s > EPS * s[0]))
    return Vt[:k].T @ ((U[:, :k].T @ b) / s[:k])


def chol_solve(G: np.ndarray, b: np.ndarray) -> np.ndarray:
    c = sla.cho_factor(G, lower=True, check_finite=False)
    return sla.cho_solve(c, b, check_finite=False)


def qr_solve(M: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Householder QR least squares (LAPACK geqrf/ormqr/trtrs): a second backward-stable solver of the SAME problem as
    ``svd_solve``; the two bracket what "the reference's answer" means on an ill-conditioned problem (they differ by
    ~cond(M)*eps).  Not a reference code path -- a comparator for tests."""
    Q, R = np.linalg.qr(M)
    return sla.solve_triangular(R, Q.T @ b, check_finite=False)


def fourier_solve(A, y, zerofreq, lam=0.0, mode="literal"):
    """src/utilities.jl:56-60: x = svd([A; lam I]) \\ [y; 0]  (ridge lam^2, Q4).  mode "qr": same problem by QR."""
    n = A.shape[1]
    if mode in ("literal", "qr"):
        solve = svd_solve if mode == "literal" else qr_solve
        if lam > 0:
            x = solve(np.vstack([A, lam * np.eye(n)]), np.concatenate([y, np.zeros(n)]))
        else:
            x = solve(A, y)
    else:
        x = chol_solve(A.T @ A + (lam * lam) * np.eye(n), A.T @ y)
    return fourier2complex(x, zerofreq)


def ls_spectral(y, t, f=None, W=None, lam=1e-10, mode="literal"):
    """src/lsfft.jl:62-67 (unweighted, ridge lam^2) and :74-80 (weighted, ridge lam, Q5)."""
    y = np.asarray(y, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    if f is None:
        f = default_freqs(t)
    f = np.asarray(f, dtype=np.float64)
    A, zerofreq = get_fourier_regressor(t, f)
    if W is None:
        return fourier_solve(A, y, zerofreq, lam, mode), f
    W = np.asarray(W, dtype=np.float64)
    AtW = A.T * W[None, :]
    M = AtW @ A + lam * np.eye(A.shape[1])
    if mode == "literal":
        # ((A'Wd*A + lam I) \ (A'Wd)) * y : LU with N right-hand sides, then a GEMV (src/lsfft.jl:77)
        lu = sla.lu_factor(M, check_finite=False)
        x = sla.lu_solve(lu, AtW, check_finite=False) @ y
    else:
        x = chol_solve(M, AtW @ y)
    return fourier2complex(x, zerofreq), f


def tls_spectral(y, t, f=None):
    """src/lsfft.jl:87-99: total least squares by the SVD of [A y] (LAPACK gesvd 'S','S'): x = -V21 / V22 with
    V21 = Vt[n+1, 1:n], V22 = Vt[n+1, n+1].  Default frequencies default_freqs(t)[1:end-1]."""
    y = np.asarray(y, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    if f is None:
        f = default_freqs(t)[:-1]
    f = np.asarray(f, dtype=np.float64)
    A, zerofreq = get_fourier_regressor(t, f)
    n = A.shape[1]
    _, _, Vt = sla.svd(np.column_stack([A, y]), full_matrices=False, lapack_driver="gesvd")
    x = -Vt[n, :n] / Vt[n, n]
    return fourier2complex(x, zerofreq), f


# --------------------------------------------------------------------------------------------------------
# windowed estimators -- src/lsfft.jl:112-193
# --------------------------------------------------------------------------------------------------------


def _abs2(z):
    return z.real * z.real + z.imag * z.imag  # Julia abs2 (Q8: not |z|**2)


def _mul_conj(a, b):
    """a .* conj.(b) in Julia's unfused order: re = fl(fl(ar*br)+fl(ai*bi)), im = fl(fl(ai*br)-fl(ar*bi)).

    numpy's complex multiply may contract to FMA, which breaks the exact ``ls_cohere(y,y,t) .== 1`` KAT."""
    re = a.real * b.real + a.imag * b.imag
    im = a.imag * b.real - a.real * b.imag
    return re + 1j * im


def ls_windowpsd(y, t, freqs=None, nw=8, noverlap=-1, window_func=rect, estimator=None, mode="literal", **kw):
    """src/lsfft.jl:112-126.  S = sum |x_i|^2 / K^2, always the weighted estimator (Q6)."""
    y = np.asarray(y, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    n = len(y) // nw
    if freqs is None:
        freqs = default_freqs(t, n)
    est = estimator or (lambda yi, ti, fr, W, **k: ls_spectral(yi, ti, fr, W, mode=mode, **k))
    win = Windows((y, t), n, noverlap, window_func)
    K = len(win)
    S = np.zeros(len(freqs))
    for yi, ti in win:
        x = est(yi, ti, freqs, win.W, **kw)[0]
        S += _abs2(x)
    return S / float(K) ** 2, freqs


def ls_windowcsd(y, u, t, freqs=None, nw=10, noverlap=-1, window_func=rect, estimator=None, mode="literal", **kw):
    """src/lsfft.jl:140-156.  S = sum xy*conj(xu) / K  (K, not K^2: Q7)."""
    y = np.asarray(y, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    n = len(y) // nw
    if freqs is None:
        freqs = default_freqs(t, n)
    est = estimator or (lambda yi, ti, fr, W, **k: ls_spectral(yi, ti, fr, W, mode=mode, **k))
    win = Windows((y, t, u), n, noverlap, window_func)
    K = len(win)
    S = np.zeros(len(freqs), dtype=np.complex128)
    for yi, ti, ui in win:
        xy = est(yi, ti, freqs, win.W, **kw)[0]
        xu = est(ui, ti, freqs, win.W, **kw)[0]
        S = S + _mul_conj(xy, xu)
    return (S.real / K) + 1j * (S.imag / K), freqs  # Julia Complex/Real is component-wise


def ls_cohere(y, u, t, freqs=None, nw=10, noverlap=-1, estimator=None, mode="literal", **kw):
    """src/lsfft.jl:176-193.  Window hard-coded to hanning (Q8)."""
    y = np.asarray(y, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    n = len(y) // nw
    if freqs is None:
        freqs = default_freqs(t, n)
    est = estimator or (lambda yi, ti, fr, W, **k: ls_spectral(yi, ti, fr, W, mode=mode, **k))
    win = Windows((y, t, u), n, noverlap, hanning)
    Syy = np.zeros(len(freqs))
    Suu = np.zeros(len(freqs))
    Syu = np.zeros(len(freqs), dtype=np.complex128)
    for yi, ti, ui in win:
        xy = est(yi, ti, freqs, win.W, **kw)[0]
        xu = est(ui, ti, freqs, win.W, **kw)[0]
        Syu += _mul_conj(xy, xu)
        Syy += _abs2(xy)
        Suu += _abs2(xu)
    return _abs2(Syu) / (Suu * Syy), freqs


# --------------------------------------------------------------------------------------------------------
# LPV (Fourier x RBF) -- src/utilities.jl:23-36, src/lsfft.jl:195-259
# --------------------------------------------------------------------------------------------------------


def basis_centers(V, Nv, coulomb):
    """Centres and gamma of basis_activation_func (src/utilities.jl:23-36)."""
    V = np.asarray(V, dtype=np.float64)
    if coulomb:
        vc = np.linspace(0.0, np.max(np.abs(V)), Nv + 2)[1:-1]
        vc = np.concatenate([-vc[::-1], vc])
        gamma = (2 * Nv) / abs(vc[0] - vc[-1])
    else:
        vc = np.linspace(np.min(V), np.max(V), Nv)
        gamma = Nv / abs(vc[0] - vc[-1])
    return vc, gamma


def basis_activation(V, Nv, normalize=True, coulomb=False):
    """K(V) evaluated for all samples at once: N x Nv (or N x 2Nv with coulomb). src/lsfft.jl:195-207."""
    V = np.asarray(V, dtype=np.float64)
    vc, gamma = basis_centers(V, Nv, coulomb)
    K = np.exp(-gamma * (V[:, None] - vc[None, :]) ** 2)
    if coulomb:
        K = K * (np.sign(V)[:, None] == np.sign(vc)[None, :])
    if normalize:
        K = K / np.sum(K, axis=1, keepdims=True)
    return K


def lpv_regressor(X, V, w, Nv, normalize=True, coulomb=False):
    """A[n, f+(k-1)Nf] = exp(-i w_f X_n) K_k(V_n) (src/lsfft.jl:244-248; conj from the trailing adjoint, Q10).

    Returns Ar = [Re A, Im A] (N x 2*Nf*Nvv)."""
    X = np.asarray(X, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64).ravel()
    K = basis_activation(V, Nv, normalize, coulomb)  # N x Nvv
    ph = np.outer(X, w)  # N x Nf   (w .* X)
    C, S = np.cos(ph), np.sin(ph)
    # column index f + k*Nf  -> (k, f) with f fastest
    re = (K[:, :, None] * C[:, None, :]).reshape(len(X), -1)
    im = (K[:, :, None] * (-S)[:, None, :]).reshape(len(X), -1)
    return np.concatenate([re, im], axis=1)


@dataclass
class SpectralExt:
    """src/LPVSpectral.jl:59-70."""

    Y: np.ndarray
    X: np.ndarray
    V: np.ndarray
    w: np.ndarray
    Nv: int
    lam: float
    coulomb: bool
    normalize: bool
    x: np.ndarray
    Sigma: Optional[np.ndarray]
    fva: Optional[float] = None


def reshape_params(x, Nf):
    """src/utilities.jl:77 -- Nf x (len/Nf), column major."""
    return np.reshape(x, (Nf, -1), order="F")


def psd(se: SpectralExt) -> np.ndarray:
    """src/lsfft.jl:214-217."""
    rp = reshape_params(se.x, len(se.w))
    s = np.sum(rp, axis=1)
    return _abs2(s)


def ls_spectral_lpv(Y, X, V, w, Nv, lam=1e-8, coulomb=False, normalize=True, mode="literal", want_sigma=True):
    """src/lsfft.jl:239-259: solve ridge lam^2 (pivoted QR on [Ar; lam I]); Sigma ridge lam (Q10)."""
    Y = np.asarray(Y, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64).ravel()
    Ar = lpv_regressor(X, V, w, Nv, normalize, coulomb)
    n2 = Ar.shape[1]
    if mode == "literal":
        if lam > 0:
            M = np.vstack([Ar, lam * np.eye(n2)])
            rhs = np.concatenate([Y, np.zeros(n2)])
            xr = sla.lstsq(M, rhs, lapack_driver="gelsy", check_finite=False)[0]
        else:
            xr = sla.lstsq(Ar, Y, lapack_driver="gelsy", check_finite=False)[0]
    else:
        xr = chol_solve(Ar.T @ Ar + lam * lam * np.eye(n2), Ar.T @ Y)
    n = n2 // 2
    params = xr[:n] + 1j * xr[n:]
    e = Ar @ xr - Y
    ve = np.var(e, ddof=1)
    Sigma = None
    if want_sigma:
        Sigma = ve * np.linalg.inv(Ar.T @ Ar + lam * np.eye(n2))
    fva = 1.0 - ve / np.var(Y, ddof=1)
    return SpectralExt(Y, np.asarray(X), np.asarray(V), w, Nv, lam, coulomb, normalize, params, Sigma, fva)


def ls_windowpsd_lpv(Y, X, V, w, Nv, nw=10, noverlap=0, mode="literal", **kw):
    """src/lsfft.jl:267-277 (rect window, S not normalised)."""
    w = np.asarray(w, dtype=np.float64).ravel()
    Y = np.asarray(Y, dtype=np.float64)
    S = np.zeros(len(w))
    win = Windows((Y, np.asarray(X, dtype=np.float64), np.asarray(V, dtype=np.float64)), len(Y) // nw, noverlap, rect)
    for y, x, v in win:
        se = ls_spectral_lpv(y, x, v, w, Nv, mode=mode, want_sigma=False, **kw)
        S += psd(se)
    return S


# --------------------------------------------------------------------------------------------------------
# proximal operators (ProximalOperators.jl, restated) and ADMM -- src/lasso.jl:136-171
# --------------------------------------------------------------------------------------------------------


@dataclass
class NormL1:
    lam: float = 1.0

    def prox(self, v, gamma):
        return np.sign(v) * np.maximum(np.abs(v) - gamma * self.lam, 0.0)

    def value(self, x):
        return self.lam * np.sum(np.abs(x))


@dataclass
class NormL0:
    lam: float = 1.0

    def prox(self, v, gamma):
        return np.where(np.abs(v) > math.sqrt(2.0 * gamma * self.lam), v, 0.0)

    def value(self, x):
        return self.lam * np.count_nonzero(x)


@dataclass
class IndBallL0:
    r: int = 1

    def prox(self, v, gamma):
        # keep the r largest |v_i|; ties broken towards the lower index (stable sort)
        order = np.argsort(-np.abs(v), kind="stable")[: self.r]
        z = np.zeros_like(v)
        z[order] = v[order]
        return z

    def value(self, x):
        return 0.0 if np.count_nonzero(x) <= self.r else math.inf


@dataclass
class GroupNormL2:
    """SlicedSeparableSum of NormL2(lam) over contiguous groups of ``gsize`` covering the first ngroups*gsize."""

    lam: float
    gsize: int
    ngroups: int

    def prox(self, v, gamma):
        z = np.zeros_like(v)  # uncovered entries are never written by SlicedSeparableSum and stay 0 (Q16)
        m = self.gsize * self.ngroups
        g = v[:m].reshape(self.ngroups, self.gsize)
        nrm = np.sqrt(np.sum(g * g, axis=1))
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = np.where(nrm > 0, np.maximum(0.0, 1.0 - gamma * self.lam / nrm), 0.0)
        z[:m] = (g * scale[:, None]).ravel()
        return z

    def value(self, x):
        m = self.gsize * self.ngroups
        g = x[:m].reshape(self.ngroups, self.gsize)
        return self.lam * np.sum(np.sqrt(np.sum(g * g, axis=1)))


def _cg(apply_op, b, x0, reltol, maxiter):
    """IterativeSolvers.cg! restated: tolerance relative to the INITIAL residual norm, warm start x0."""
    x = x0.copy()
    r = b - apply_op(x)
    p = r.copy()
    rs = r @ r
    tol = reltol * math.sqrt(rs)
    it = 0
    while it < maxiter and math.sqrt(rs) > tol:
        Ap = apply_op(p)
        alpha = rs / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        rs_new = r @ r
        p = r + (rs_new / rs) * p
        rs = rs_new
        it += 1
    return x, it


class QuadProx:
    """x-update operators.

    kind="ls"   : LeastSquares(A,y,iterative=true): f=0.5||Ax-y||^2, prox solves (G + I/mu) x = A'y + v/mu
    kind="quad" : Quadratic(Q,q,iterative=true):   f=0.5x'Qx + q'x, prox solves (Q + I/mu) x = v/mu - q  (Q13)

    Built from the Gram matrix only (the reference's Tall branch caches S=A'A as well).  ``mode="literal"``
    uses the warm-started CG (warm start = the prox input, as ProximalOperators does ``y .= x`` first);
    ``mode="gram"`` factorises (G + I/mu) once and solves exactly.
    """

    def __init__(self, G, b, kind="ls", mode="gram"):
        self.G = G
        self.b = b
        self.kind = kind
        self.mode = mode
        self._fact = None
        self._mu = None
        self.cg_its = 0

    def rhs(self, v, mu):
        return (self.b + v / mu) if self.kind == "ls" else (v / mu - self.b)

    def prox(self, v, mu):
        n = len(v)
        if self.mode == "literal":
            op = lambda p: self.G @ p + p / mu
            x, it = _cg(op, self.rhs(v, mu), v, math.sqrt(EPS), n)
            self.cg_its += it
            return x
        if self._fact is None or self._mu != mu:
            self._fact = sla.cho_factor(self.G + np.eye(n) / mu, lower=True, check_finite=False)
            self._mu = mu
        return sla.cho_solve(self._fact, self.rhs(v, mu), check_finite=False)


def admm(x0, proxf: QuadProx, proxg, iters=10000, tol=1e-5, mu=0.05, printerval=100, cb=None, log=None):
    """src/lasso.jl:136-171.  Returns (x, z, iterations_done, last ||x-z||)."""
    assert 0 <= mu <= 1, "mu should be <= 1"
    x = np.array(x0, dtype=np.float64)
    z = x.copy()
    u = np.zeros_like(x)
    nxz = math.inf
    done = 0
    for i in range(1, iters + 1):
        x = proxf.prox(z - u, mu)
        z = proxg.prox(x + u, mu)
        tmp = x - z
        u = u + tmp
        nxz = float(np.sqrt(tmp @ tmp))
        done = i
        if i % printerval == 0:
            if log is not None:
                log.append((i, nxz))
            if cb is not None:
                cb(x, z)
        if nxz < tol:
            if log is not None:
                log.append((i, nxz))
            break
    return x, z, done, nxz


def ls_sparse_spectral(y, t, f=None, W=None, init=False, lam=1.0, proxg=None, mode="gram", return_info=False,
                       **kw):
    """src/lasso.jl:85-102 (unweighted) and :105-126 (weighted, sign quirk Q13)."""
    y = np.asarray(y, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    if f is None:
        f = default_freqs(t)
    f = np.asarray(f, dtype=np.float64)
    if proxg is None:
        proxg = NormL1(lam)
    A, zerofreq = get_fourier_regressor(t, f)
    if init:
        params = fourier_solve(A, y, zerofreq, lam, "literal" if mode == "literal" else "gram")  # Q14
    else:
        params = np.zeros(len(f), dtype=np.complex128)
    x0 = complex2fourier(params, zerofreq)
    if W is None:
        pf = QuadProx(A.T @ A, A.T @ y, "ls", mode)
    else:
        W = np.asarray(W, dtype=np.float64)
        AtW = A.T * W[None, :]
        pf = QuadProx(AtW @ A, AtW @ y, "quad", mode)
    x, z, its, nxz = admm(x0, pf, proxg, **kw)
    out = fourier2complex(z, zerofreq)
    if return_info:
        return out, f, dict(iters=its, residual=nxz, x=x, z=z, A=A)
    return out, f


def lpv_group_perm(Nf, ncols2):
    """inds of src/lasso.jl:47: permuted position p=(f)*2Nv+k  <->  original f+k*Nf (0-based)."""
    return np.arange(ncols2).reshape(-1, Nf).T.ravel()


def ls_sparse_spectral_lpv(y, X, V, w, Nv, lam=1.0, coulomb=False, normalize=True, mode="gram",
                           return_info=False, **kw):
    """src/lasso.jl:27-70 (group lasso over frequencies)."""
    y = np.asarray(y, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64).ravel()
    Nf = len(w)
    Ar = lpv_regressor(X, V, w, Nv, normalize, coulomb)
    inds = lpv_group_perm(Nf, Ar.shape[1])
    Phi = Ar[:, inds]
    pf = QuadProx(Phi.T @ Phi, Phi.T @ y, "ls", mode)
    pg = GroupNormL2(lam, 2 * Nv, Nf)
    x0 = np.zeros(Phi.shape[1])
    x, z, its, nxz = admm(x0, pf, pg, **kw)
    zs = z[np.argsort(inds, kind="stable")]
    h = len(zs) // 2
    params = zs[:h] + 1j * zs[h:]
    se = SpectralExt(y, np.asarray(X), np.asarray(V), w, Nv, lam, coulomb, normalize, params, None)
    if return_info:
        return se, dict(iters=its, residual=nxz, x=x, z=z, Phi=Phi, proxg=pg)
    return se


def sparse_objective(A, y, z, proxg):
    """0.5||Az-y||^2 + g(z): the quantity the 1e-8 ADMM parity bar is stated on."""
    r = A @ z - y
    return 0.5 * float(r @ r) + float(proxg.value(z))


# --------------------------------------------------------------------------------------------------------
# synthetic signals (numpy restatement of the reference's generator, test/runtests.jl:6-17)
# --------------------------------------------------------------------------------------------------------


def generate_lpv_signal(N, seed=0, modphase=True):
    rng = np.random.default_rng(seed)
    x = np.sort(10.0 * rng.random(N))
    v = np.linspace(0.0, 1.0, N)
    fn = [lambda v: 2 * v ** 2, lambda v: 2 / (5 * v + 1), lambda v: 3 * np.exp(-10 * (v - 0.5) ** 2)]
    w = 2 * np.pi * np.array([2.0, 10.0, 20.0])
    dep = np.stack([fn[i % 3](v) for i in range(len(w))], axis=1)
    fm = np.cos(x[:, None] * w[None, :] - 0.5 * modphase * dep)
    y = np.sum(dep * fm, axis=1) + 0.1 * rng.standard_normal(N)
    return y, v, x
