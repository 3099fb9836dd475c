#!/bin/bash
# where do the diagonal Gram tiles spend their time?  (1) k_gram durations with diagonal / off-diagonal tiles skipped,
# (2) source-level ncu capture of the production kernel, (3) the C-ABI tests
set -u
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-extra --no-admm"
for v in skipdiag skipoff; do
  LPVS_LIB=$PWD/gpurun_variants/$v.so ncu --clock-control none --metrics gpu__time_duration.sum -k "regex:^k_gram$" -c 3 --csv --log-file $OUT/dur_${v}_$TAG.csv $BENCH > $OUT/dur_${v}_$TAG.log 2>&1
  echo $v; grep k_gram $OUT/dur_${v}_$TAG.csv | awk -F'","' '{print $NF}' | head -3
done
ncu --clock-control none --metrics gpu__time_duration.sum -k "regex:^k_gram$" -c 3 --csv --log-file $OUT/dur_full_$TAG.csv $BENCH > $OUT/dur_full_$TAG.log 2>&1
echo full; grep k_gram $OUT/dur_full_$TAG.csv | awk -F'","' '{print $NF}' | head -3
ncu --clock-control none --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gram_$TAG -f $BENCH > $OUT/ncu_gram_$TAG.log 2>&1
ncu -i $OUT/gram_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/gram_${TAG}_source.csv 2>/dev/null
ls -la $OUT/gram_${TAG}_source.csv $OUT/gram_$TAG.ncu-rep
rm -f $OUT/gram_$TAG.ncu-rep
timeout 600 python -m pytest tests/test_gpu_cabi.py -m gpu -x -q 2>&1 | tail -3
