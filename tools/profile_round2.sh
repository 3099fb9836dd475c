#!/bin/bash
# Round-2 profiling pass on the GPU box (1 GPU): plain runs first (must exit 0), then the ncu launch list of the same
# bench command, then one --set full capture of the Gram kernel (default reference-phase mode) and of the QR path's GEMM.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh r02'
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-extra"
$BENCH > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; exit 1; }
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
$NCU --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gram_$TAG -f $BENCH --no-admm > $OUT/ncu_gram_$TAG.log 2>&1
python tools/measure_configs.py cfg1 > $OUT/plain_cfg1_$TAG.log 2>&1 || { echo "plain cfg1 failed"; exit 1; }
$NCU --set full --import-source on -k regex:k_gemm_nt -s 2 -c 2 -o $OUT/gemmnt_$TAG -f python tools/measure_configs.py cfg1 > $OUT/ncu_gemmnt_$TAG.log 2>&1
for r in gram gemmnt; do
    ncu -i $OUT/${r}_$TAG.ncu-rep --page raw --csv > $OUT/${r}_${TAG}_raw.csv 2>/dev/null
done
tail -c 400 $OUT/plain_$TAG.json
