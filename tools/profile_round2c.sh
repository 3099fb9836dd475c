#!/bin/bash
# potf2 probe, cfg1 with / without the critical-path priority stream, GPU tests, short bench
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
tools/potf2_probe 2047 > $OUT/potf2_$TAG.log 2>&1; cat $OUT/potf2_$TAG.log
python tools/measure_configs.py cfg1 > $OUT/cfg1_crit_$TAG.log 2>&1; cp $OUT/measure_cfg1.json $OUT/measure_cfg1_crit_$TAG.json
LPVS_NO_CRIT_STREAM=1 python tools/measure_configs.py cfg1 > $OUT/cfg1_nocrit_$TAG.log 2>&1; cp $OUT/measure_cfg1.json $OUT/measure_cfg1_nocrit_$TAG.json
grep -h call_ms $OUT/cfg1_crit_$TAG.log $OUT/cfg1_nocrit_$TAG.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 --cpu-windows 2 --no-admm > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
python - <<'P'
import json
d=json.loads(open("gpurun_out/plain_r02c.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print("cfg1", d["extra"]["cfg1"]["ms_per_spectrum"])
P
