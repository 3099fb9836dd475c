#!/bin/bash
# balanced diagonal-piece deal: Gram tests, bench, diag-only kernel duration
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fourier.py tests/test_gpu_baseline_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
BENCH="python bench.py --steps 5 --warmup 3 --cpu-windows 2 --no-extra --no-admm"
$BENCH > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/plain_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
P
LPVS_LIB=$PWD/gpurun_variants/skipoff.so ncu --clock-control none --metrics gpu__time_duration.sum -k "regex:^k_gram$" -c 3 --csv --log-file $OUT/dur_skipoff_$TAG.csv $BENCH > $OUT/dur_skipoff_$TAG.log 2>&1
echo skipoff; grep k_gram $OUT/dur_skipoff_$TAG.csv | awk -F'","' '{print $NF}' | head -3
