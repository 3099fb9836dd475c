"""world_size-2 gloo tests of the multi-GPU host logic (window sharding + accumulator all-reduce; row-sharded
Gram all-reduce), with the oracle standing in for the per-rank GPU compute.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import lpvs_oracle as o


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_sums(kind, y, u, t, f, W, n, noverlap, lam, k0, k1):
    hop = n - noverlap
    nf = len(f)
    syy, suu = np.zeros(nf), np.zeros(nf)
    syu = np.zeros(nf, dtype=np.complex128)
    for k in range(k0, k1):
        sl = slice(k * hop, k * hop + n)
        xy, _ = o.ls_spectral(y[sl], t[sl], f, W, lam=lam, mode="gram")
        syy += o._abs2(xy)
        if u is not None:
            xu, _ = o.ls_spectral(u[sl], t[sl], f, W, lam=lam, mode="gram")
            suu += o._abs2(xu)
            syu += o._mul_conj(xy, xu)
    if kind == 0:
        return syy
    if kind == 1:
        return np.concatenate([syu.real, syu.imag])
    return np.concatenate([syy, suu, syu.real, syu.imag])


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lpvspectral_jl_b200 import _dist as D

    rng = np.random.default_rng(5)
    N, nw = 2400, 12
    t = np.sort(10 * rng.random(N))
    y = np.sin(2 * np.pi * 9 * t) + 0.2 * rng.standard_normal(N)
    u = 0.6 * y + 0.3 * rng.standard_normal(N)
    n = N // nw
    f = np.arange(12) * 2.0 / (t[n] - t[0])
    W = o.hanning(n)
    out = {}
    for kind in (0, 1, 2):
        res, K = D.ls_window_sharded(kind, y, u if kind else None, t, f, n=n, W=W, lam=1e-10, sums_fn=_oracle_sums)
        out[kind] = (res, K)
    # row-sharded Gram: partial sums over contiguous row blocks add up to the full Gram
    r0, r1 = D.shard_range(N, rank, world)
    A, _ = o.get_fourier_regressor(t[r0:r1], f)
    part = np.concatenate([(A.T @ A).ravel(), A.T @ y[r0:r1]])
    tot = D._allreduce_sum_np(part)
    out["gram"] = tot
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_window_and_row_sharding_world2():
    from lpvspectral_jl_b200 import _dist as D

    assert D.shard_range(2047, 0, 8) == (0, 255) and D.shard_range(2047, 7, 8) == (1791, 2047)
    covered = []
    for r in range(8):
        a, b = D.shard_range(2047, r, 8)
        covered += list(range(a, b))
    assert covered == list(range(2047))

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    rng = np.random.default_rng(5)
    N, nw = 2400, 12
    t = np.sort(10 * rng.random(N))
    y = np.sin(2 * np.pi * 9 * t) + 0.2 * rng.standard_normal(N)
    u = 0.6 * y + 0.3 * rng.standard_normal(N)
    n = N // nw
    f = np.arange(12) * 2.0 / (t[n] - t[0])
    S, _ = o.ls_windowpsd(y, t, f, nw=nw, window_func=o.hanning, mode="gram")
    Cs, _ = o.ls_windowcsd(y, u, t, f, nw=nw, window_func=o.hanning, mode="gram")
    Co, _ = o.ls_cohere(y, u, t, f, nw=nw, mode="gram")
    assert out[0][1] == 23
    assert np.allclose(out[0][0], S, rtol=1e-12, atol=0)
    assert np.allclose(out[1][0], Cs, rtol=1e-11, atol=1e-14)
    assert np.allclose(out[2][0], Co, rtol=1e-11, atol=0)
    A, _ = o.get_fourier_regressor(t, f)
    full = np.concatenate([(A.T @ A).ravel(), A.T @ y])
    assert np.allclose(out["gram"], full, rtol=1e-12, atol=1e-12)


def _failing_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lpvspectral_jl_b200 import _dist as D

    rng = np.random.default_rng(6)
    N, n = 1200, 100
    t = np.sort(10 * rng.random(N))
    y = rng.standard_normal(N)
    f = np.arange(4) * 2.0 / (t[n] - t[0])

    def sums(kind, yy, uu, tt, ff, W, nn, nov, lam, k0, k1):
        if rank == 1:
            raise ValueError("rank-local failure (stands in for NOT_SPD / NONFINITE / NOMEM in one window range)")
        return _oracle_sums(kind, yy, uu, tt, ff, W, nn, nov, lam, k0, k1)

    try:
        D.ls_window_sharded(0, y, None, t, f, n=n, W=o.hanning(n), lam=1e-10, sums_fn=sums)
        q.put((rank, "returned"))
    except ValueError:
        q.put((rank, "own"))
    except RuntimeError as e:
        q.put((rank, "peer" if "another rank" in str(e) else "other"))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_local_failure_raises_on_every_rank():
    """A liblpvs error on one rank used to leave the others blocked in the accumulator all-reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: "peer", 1: "own"}
