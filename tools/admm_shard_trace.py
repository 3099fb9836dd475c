"""torchrun --nproc-per-node P tools/admm_shard_trace.py : phase timeline of the sharded ADMM loop (rank 0), cfg3."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
local = int(os.environ.get("LOCAL_RANK", "0"))
TRACE = os.path.join(ROOT, "gpurun_out", f"shard_trace_{local}.bin")
os.makedirs(os.path.dirname(TRACE), exist_ok=True)
import bench  # noqa: E402
import lpvspectral_jl_b200 as lp  # noqa: E402
from lpvspectral_jl_b200 import _dist as D, _lib as L  # noqa: E402

torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = lp.Context(local)
t, y, f = bench.make_cfg3()
h = C.c_void_p()
ctx.check(ctx.lib.lpvs_admm_create_fourier(ctx.h, y.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), len(y),
                                           f.ctypes.data_as(C.c_void_p), len(f), None, L.PROX_L1, 0.1, 0.05, None, 0, 0.0,
                                           C.byref(h)))
s = D.admm_shard(lp.ADMM(ctx, h))
s.step(100, 0.0)
dist.barrier()
os.environ["LPVS_ADMM_TRACE"] = TRACE
s.step(400, 0.0)
ms, _ = s.timing()
dist.barrier()
s.free()
if rank == 0:
    raw = open(TRACE, "rb").read()
    ni, grid = np.frombuffer(raw[:8], dtype=np.int32)
    tr = np.frombuffer(raw[8:], dtype=np.int64).reshape(ni, grid, 8).astype(np.float64) / 1965.0
    names = ["phase1+publish", "barrier1", "phase2a scatter+fence", "barrier2+signal A", "wait A", "phase2b+barrier3",
             "signal B + wait B"]
    print(f"world {world}: {400 / ms * 1e3:.0f} it/s = {ms / 400 * 1e3:.1f} us/it")
    d = np.diff(tr[1], axis=1)
    for k, nm in enumerate(names):
        print(f"   {nm:24s} min {d[:, k].min():7.2f}  med {np.median(d[:, k]):7.2f}  max {d[:, k].max():7.2f}")
    print("   start-to-start", np.median(tr[2, :, 0] - tr[1, :, 0]))
dist.destroy_process_group()
