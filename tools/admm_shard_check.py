"""torchrun --nproc-per-node P tools/admm_shard_check.py [full]: ONE L1 ADMM problem sharded over P GPUs (device-initiated
peer stores over NVLink) vs the same problem on one GPU: iterates, iteration count, iterations/s.
`full` = BASELINE cfg3 (16 383 unknowns); default = a 4 095-unknown problem (quick); `ball` = the same quick problem with the
IndBallL0 prox (keep the 24 largest); `lpv` / `lpvfull` = the group lasso of ls_sparse_spectral_lpv."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _dist as D, _lib as L  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = lp.Context(local)
if "LPVS_SHARD_EXCHANGE" in os.environ:  # 1 (default): per-CTA arrival counters; 0: grid barrier + flag hops
    ctx.set_option(L.OPT_SHARD_EXCHANGE, int(os.environ["LPVS_SHARD_EXCHANGE"]))
full = len(sys.argv) > 1 and sys.argv[1] == "full"
ball = len(sys.argv) > 1 and sys.argv[1] == "ball"
lpv = len(sys.argv) > 1 and sys.argv[1] in ("lpv", "lpvfull")  # group lasso (BASELINE cfg4 with "lpvfull")
if lpv:
    from oracle import lpvs_oracle as o

    big = sys.argv[1] == "lpvfull"
    Y, V, X = o.generate_lpv_signal(20000 if big else 4000, seed=4)
    w = 2 * np.pi * np.arange(1, (64 if big else 16) + 1) * 0.4
    Nv = 50 if big else 20
elif full:
    t, y, f = bench.make_cfg3()
else:
    rng = np.random.default_rng(3)
    N = 4096
    t = np.sort(10 * rng.random(N))
    f = lp.default_freqs(t)[:2048]
    y = sum(np.cos(2 * np.pi * f[k] * t + k) for k in (100, 500, 900)) + 0.1 * rng.standard_normal(N)


def create():
    h = C.c_void_p()
    if lpv:
        vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        ctx.check(ctx.lib.lpvs_admm_create_lpv(ctx.h, vp(Y), vp(X), vp(V), len(Y), vp(w), len(w), Nv, 0, 1, 0.1, 0.05,
                                               C.byref(h)))
        return lp.ADMM(ctx, h)
    ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, y.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                               len(y), f.ctypes.data_as(C.c_void_p), len(f), None,
                                               L.PROX_BALL_L0 if ball else L.PROX_L1, 24.0 if ball else 0.1, 0.05, None, 0,
                                               0.0, C.byref(h)))
    return lp.ADMM(ctx, h)


iters = 400 if full else 600
tol_run = 1e-9 if not lpv else 1e-7
one = create()
one.step(iters, tol_run)
x1, z1 = one.get()
it1 = one.iters
one.step(iters, 0.0)
ms1, _ = one.timing()
one.free()
sh = D.admm_shard(create())
sh.step(iters, tol_run)
dist.barrier()
xs, zs = sh.get()
its = sh.iters
sh.step(iters, 0.0)
mss, _ = sh.timing()
dist.barrier()
sh.free()
err = np.linalg.norm(zs - z1) / max(np.linalg.norm(z1), 1e-300)
errx = np.linalg.norm(xs - x1) / np.linalg.norm(x1)
supp = bool(np.array_equal(zs != 0, z1 != 0))
ok = err < 1e-9 and errx < 1e-9 and abs(its - it1) <= 1 and supp
flag = torch.tensor([1.0 if ok else 0.0], device=torch.device("cuda", local))
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
t_all = torch.tensor([mss], device=torch.device("cuda", local))
dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": ("cfg4_group_lasso" if big else "group_lasso_640") if lpv else
                      ("cfg3_l1_admm" if full else ("ball_l0_admm_4095" if ball else "l1_admm_4095")), "n_gpus": world, "iters": iters,
                      "one_gpu_it_per_s": iters / ms1 * 1e3, "sharded_it_per_s": iters / t_all.item() * 1e3,
                      "speedup": ms1 / t_all.item(), "rel_err_z": err, "rel_err_x": errx, "iters_one": it1,
                      "iters_sharded": its, "same_support": supp, "ok": bool(flag.item() == 1.0)}))
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
