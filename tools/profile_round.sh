#!/bin/bash
# One profiling pass on the GPU box (run through gpurun, 1 GPU): plain runs first (must exit 0), then the ncu launch
# list of the same bench command, then one --set full capture per dominant kernel.  Outputs land in gpurun_out/;
# the summaries that are judged get copied to profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r01c'
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python bench.py --steps 2 --warmup 1 --cpu-windows 2 > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; exit 1; }
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --cpu-windows 2 > $OUT/ncu_launches_$TAG.log 2>&1
$NCU --set full --import-source on -k "regex:^k_gram$" -s 2 -c 1 -o $OUT/gram_$TAG -f \
    python bench.py --steps 2 --warmup 1 --cpu-windows 2 --no-admm > $OUT/ncu_gram_$TAG.log 2>&1
python tools/admm_run.py 200 cfg3 > $OUT/plain_admm3_$TAG.log 2>&1 || { echo "plain admm cfg3 failed"; exit 1; }
$NCU --set full --import-source on -k regex:k_admm_symv -c 1 -o $OUT/admm3_$TAG -f \
    python tools/admm_run.py 200 cfg3 > $OUT/ncu_admm3_$TAG.log 2>&1
python tools/admm_run.py 1000 cfg4 > $OUT/plain_admm4_$TAG.log 2>&1 || { echo "plain admm cfg4 failed"; exit 1; }
$NCU --set full --import-source on -k regex:k_admm_symv -c 1 -o $OUT/admm4_$TAG -f \
    python tools/admm_run.py 1000 cfg4 > $OUT/ncu_admm4_$TAG.log 2>&1
for r in gram admm3 admm4; do
    ncu -i $OUT/${r}_$TAG.ncu-rep --page raw --csv > $OUT/${r}_${TAG}_raw.csv 2>/dev/null
done
cat $OUT/plain_admm3_$TAG.log $OUT/plain_admm4_$TAG.log
tail -c 600 $OUT/plain_$TAG.json
