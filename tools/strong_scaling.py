"""Strong scaling of BASELINE cfg2 (ONE 2^22-sample record, K=2047 windows split over the ranks).

  python -m torch.distributed.run --nproc-per-node N tools/strong_scaling.py

Every rank keeps the record resident, takes windows [K*r/P, K*(r+1)/P), and the Nf accumulators are all-reduced.
Timing: CUDA events on the shared stream around `steps` full passes, barrier + synchronize both sides, max over
ranks.  Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _dist as D, _lib as L  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
ctx = lp.Context(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
t, y, f, n = bench.make_cfg2()
K = lp.window_count(len(y), n, -1)
k0, k1 = D.shard_range(K, rank, world)
hop = n >> 1
s0, s1 = k0 * hop, (k1 - 1) * hop + n
d_t = torch.from_numpy(t[s0:s1]).cuda()
d_y = torch.from_numpy(y[s0:s1]).cuda()
W = lp.hanning(n)
sums = np.zeros(len(f))
acc = torch.zeros(len(f), dtype=torch.float64, device="cuda")
info = C.c_int(0)
vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def step():
    ctx.check(ctx.lib.lpvs_ls_window_sums_dev(ctx.h, L.WIN_PSD, C.c_void_p(d_y.data_ptr()), None,
                                              C.c_void_p(d_t.data_ptr()), s1 - s0, vp(f), len(f), vp(W), n, hop,
                                              1e-10, 0, k1 - k0, vp(sums), C.byref(info)))
    acc.copy_(torch.from_numpy(sums))
    if world > 1:
        dist.all_reduce(acc)
    return lp.window_finalize(L.WIN_PSD, acc.cpu().numpy(), len(f), K)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


steps, warm = 10, 3
for _ in range(warm):
    S = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
e0.record()
for _ in range(steps):
    S = step()
e1.record()
barrier()
wall = time.perf_counter() - w0
ms = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    per = ms.item() / steps
    print(json.dumps({"workload": "cfg2_windowpsd", "scaling": "strong", "n_gpus": world, "windows": K,
                      "ms_per_pass": per, "windows_per_s": K / per * 1e3, "peaks": np.argsort(-S)[:2].tolist()}))
if world > 1:
    dist.destroy_process_group()
