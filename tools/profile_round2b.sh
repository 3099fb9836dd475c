#!/bin/bash
# Round-2 follow-up pass (1 GPU): GPU tests, a short bench, then one --set full capture of the Gram kernel.
#   gpurun --timeout 1200 -- 'bash tools/profile_round2b.sh r02b'
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
BENCH="python bench.py --steps 5 --warmup 3 --cpu-windows 2 --no-extra --no-admm"
$BENCH > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --clock-control none --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gram_$TAG -f $BENCH > $OUT/ncu_gram_$TAG.log 2>&1
ncu -i $OUT/gram_$TAG.ncu-rep --page raw --csv > $OUT/gram_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/gram_$TAG.ncu-rep
python - <<'P'
import json,sys
d=json.loads(open("gpurun_out/plain_%s.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/plain_r02b.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
P
