"""FP64 DGEMM peak probe (cuBLAS through torch.matmul), same method as MEASURED_PEAKS.json's bf16 line.

Bench utility only -- the library never calls cuBLAS.  Writes gpurun_out/fp64_peak.json.
"""
import json
import os
import sys
import time

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dev = torch.device("cuda:0")
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    c = torch.empty(n, n, device=dev, dtype=torch.float64)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    # sustained: back to back for ~3 s
    t0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 0
    while time.time() - t0 < 3.0:
        for _ in range(4):
            torch.matmul(a, b, out=c)
            reps += 1
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # syrk-shaped and potrf timings for context
    g = a @ a.T + n * torch.eye(n, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    e0.record()
    torch.linalg.cholesky(g)
    e1.record()
    e1.synchronize()
    potrf_ms = e0.elapsed_time(e1)
    e0.record()
    torch.linalg.cholesky(g)
    e1.record()
    e1.synchronize()
    potrf_ms = min(potrf_ms, e0.elapsed_time(e1))
    out = {
        "n": n,
        "fp64_dgemm_tflops_burst": burst,
        "fp64_dgemm_tflops_sustained": sustained,
        "dgemm_ms_best": best,
        "cusolver_potrf_ms": potrf_ms,
        "potrf_tflops": n ** 3 / 3 / (potrf_ms * 1e-3) / 1e12,
        "gpu": torch.cuda.get_device_name(0),
        "how": "torch.matmul float64 n^3 (2*n^3 flop), best of 10 (burst) and back to back for 3 s (sustained)",
    }
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fp64_peak.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
