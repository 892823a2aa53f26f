#!/bin/bash
# last check of a build: GPU parity tests, per-kernel times of the default workload, one short bench line
set -u
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $OUT/${TAG}_status.txt
tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 60 python tools/profile_target.py --workload 8k1024 --frames 3 > $OUT/${TAG}_target_8k1024.log 2>&1; cat $OUT/${TAG}_target_8k1024.log
timeout 90 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_bench_8k1024_n1.json 2> $OUT/${TAG}_bench_8k1024_n1.err
echo "bench rc=$?" | tee -a $OUT/${TAG}_status.txt
head -c 300 $OUT/${TAG}_bench_8k1024_n1.json; echo
