#!/bin/bash
# bench each kernel variant in gpurun_variants/ named on the command line (plus the production build): Gram ms per step
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 5 --warmup 3 --cpu-windows 2 --no-extra --no-admm"
for v in prod "$@"; do
  if [ $v = prod ]; then unset LPVS_LIB; else export LPVS_LIB=$PWD/gpurun_variants/$v.so; fi
  $BENCH > $OUT/variant_$v.json 2> $OUT/variant_$v.err || { echo "$v failed"; tail -3 $OUT/variant_$v.err; continue; }
  python - $v <<'P'
import json,sys
d=json.loads(open("gpurun_out/variant_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step %.2f gram %.2f frac %.4f" % (d["ms_per_step"], d["roofline"]["gram_ms_per_step"], d["roofline"]["frac"]))
P
done
