#!/bin/bash
# full ncu capture of the diagonal-tiles-only variant
set -u
TAG=${1:-r02h}
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-extra --no-admm"
LPVS_LIB=$PWD/gpurun_variants/skipoff.so ncu --clock-control none --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gramdiag_$TAG -f $BENCH > $OUT/ncu_gramdiag_$TAG.log 2>&1
ncu -i $OUT/gramdiag_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/gramdiag_${TAG}_source.csv 2>/dev/null
ncu -i $OUT/gramdiag_$TAG.ncu-rep --page raw --csv > $OUT/gramdiag_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/gramdiag_$TAG.ncu-rep
ls -la $OUT/gramdiag_${TAG}_*.csv
