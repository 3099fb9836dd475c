#!/bin/bash
# GPU tests + short bench (1 GPU)
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python bench.py --steps 8 --warmup 3 --cpu-windows 2 --no-admm > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/plain_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
x=d["extra"]
print("cfg1", x["cfg1"]["ms_per_spectrum"], "exact", x["exact_phase_mode"]["gram_ms_per_step"], "cfg5a", x["cfg5a"]["s_per_pass"], x["cfg5a"]["gram_frac"], "cfg5b", x["cfg5b_rowsharded"]["gram_kernel_frac_per_gpu"], "cfg4 gram", x["cfg4"]["gram_tflops"])
P
