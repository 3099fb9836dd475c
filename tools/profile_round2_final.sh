#!/bin/bash
# Final round-2 pass (1 GPU): GPU tests, the default bench line (what the driver runs), the ncu launch list of a short bench
# command and one --set full capture of the Gram kernel.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2_final.sh r02z'
set -u
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; exit 1; }
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "gram", d["roofline"]["gram_ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
x=d["extra"]
print("cfg1", x["cfg1"]["ms_per_spectrum"], "exact", x["exact_phase_mode"]["gram_ms_per_step"], "cfg5a", x["cfg5a"]["s_per_pass"], x["cfg5a"]["gram_frac"], "cfg5b", x["cfg5b_rowsharded"]["gram_kernel_frac_per_gpu"], "cfg4", x["cfg4"]["iters_per_s"], "admm", x["admm"]["iters_per_s"])
P
SHORT="python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-extra"
ncu --clock-control none --metrics gpu__time_duration.sum -c 800 --csv --log-file $OUT/launches_$TAG.csv $SHORT > $OUT/ncu_launches_$TAG.log 2>&1
ncu --clock-control none --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gram_$TAG -f $SHORT --no-admm > $OUT/ncu_gram_$TAG.log 2>&1
ncu -i $OUT/gram_$TAG.ncu-rep --page raw --csv > $OUT/gram_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/gram_$TAG.ncu-rep
