#!/bin/bash
# Final round-2 pass (1 GPU) with LPVS_PHASE_STRUCTURED_REF in the tree: GPU tests, the default bench line (what the driver
# runs), the ncu launch list of a short bench command, the launch list + one --set full capture of k_gram_corr.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2_final2.sh r02f'
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; }
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
x=d["extra"]
print("structured", x["structured_mode"].get("ms_per_step"), "structured_ref", json.dumps(x["structured_ref_mode"])[:260])
print("cfg1", x["cfg1"]["ms_per_spectrum"], "cfg5a", x["cfg5a"]["s_per_pass"], "cfg4", x["cfg4"]["iters_per_s"], "admm", x["admm"]["iters_per_s"])
P
SHORT="python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-extra --no-admm"
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -c 800 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launches_$TAG.log 2>&1
bash tools/corr_profile.sh ${TAG}c
ncu -i $OUT/${TAG}c_gram_corr.ncu-rep --page raw --csv > $OUT/${TAG}c_gram_corr_raw.csv 2>/dev/null
