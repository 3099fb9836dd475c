"""torchrun --nproc-per-node N tools/dist_check.py : multi-GPU parity of the two sharded paths over NCCL.
Every rank checks window-sharded PSD/coherence and the row-sharded LS solve against its own single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _dist as D, _lib as L  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = lp.Context(local)
rng = np.random.default_rng(0)
N = 1 << 18
t = np.sort(10 * rng.random(N))
y = np.sin(2 * np.pi * 300 * t) + 0.3 * rng.standard_normal(N)
u = 0.7 * np.roll(y, 2) + 0.5 * rng.standard_normal(N)
n = 2048
f = np.arange(64) * 2.0 / (t[n] - t[0])
W = lp.hanning(n)
dev = torch.device("cuda", local)
ok = True
for kind, uu in ((L.WIN_PSD, None), (L.WIN_COHERE, u)):
    res, K = D.ls_window_sharded(kind, y, uu, t, f, n=n, W=W, lam=1e-10, ctx=ctx, reduce_device=dev)
    if kind == L.WIN_PSD:
        ref, _ = lp.ls_windowpsd(y, t, f, nw=N // n, window_func=lp.hanning, ctx=ctx)
    else:
        ref, _ = lp.ls_cohere(y, u, t, f, nw=N // n, ctx=ctx)
    err = np.linalg.norm(res - ref) / np.linalg.norm(ref)
    ok &= err < 1e-12
    if rank == 0:
        print(f"window-sharded kind={kind} K={K} world={world} rel err vs 1 GPU {err:.2e}")
f2 = np.arange(128) * 40.0
Wn = 0.5 + rng.random(N)
xs = D.ls_spectral_rowsharded(y, t, f2, Wn, u=u, lam=1e-10, ctx=ctx)
x1, _ = lp.ls_spectral(y, t, f2, Wn, ctx=ctx)
err = np.linalg.norm(xs[0] - x1) / np.linalg.norm(x1)
ok &= err < 1e-11
if rank == 0:
    print(f"row-sharded Gram + NCCL all-reduce world={world} rel err vs 1 GPU {err:.2e}")
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "OK" if flag.item() == 1.0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
