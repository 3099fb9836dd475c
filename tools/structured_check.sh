#!/bin/bash
# LPVS_PHASE_STRUCTURED: GPU tests + the bench line's structured_mode leg
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_structured.py tests/test_gpu_fourier.py -m gpu -x -q > gpurun_out/pytest_structured.log 2>&1; echo "pytest rc $?" >> gpurun_out/pytest_structured.log
tail -25 gpurun_out/pytest_structured.log
python bench.py --steps 3 --warmup 3 --cpu-windows 2 --no-extra --no-admm > gpurun_out/bench_structured.json 2> gpurun_out/bench_structured.err || { echo bench failed; tail -5 gpurun_out/bench_structured.err; }
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_structured.json").read().strip().splitlines()[-1])
print("headline ms/step", d["ms_per_step"], "structured", json.dumps(d["extra"].get("structured_mode")))
P
bash tools/structured_profile.sh
