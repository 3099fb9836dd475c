#!/bin/bash
# two-GPU validation: torchrun-based tests of the sharded paths, then the bench line at N = 2 (with extra.parity)
set -u
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/pytest_n2_$TAG.log 2>&1; echo "pytest rc $?" >> $OUT/pytest_n2_$TAG.log
tail -3 $OUT/pytest_n2_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err || { echo "bench N=2 failed"; tail -8 $OUT/bench_n2_$TAG.err; exit 1; }
python - $TAG <<'P'
import json,sys
d=json.loads(open("gpurun_out/bench_n2_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"])
x=d["extra"]
print("parity", json.dumps(x.get("parity")))
print("strong", json.dumps(x.get("strong"))[:300])
print("admm sharded", json.dumps(x["admm"].get("sharded"))[:400])
P
