#!/bin/bash
# One gpurun call (1 GPU) per kernel iteration: GPU parity tests, the default bench line, and the per-rank emulation of an
# 8-rank job with a few frame-batching settings.   gpurun --timeout 1200 -- 'bash tools/r02_iter.sh TAG'
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee $OUT/${TAG}_status.txt
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/profile_target.py --workload 8k1024 --frames 3 > $OUT/${TAG}_target_8k1024.log 2>&1; cat $OUT/${TAG}_target_8k1024.log
timeout 300 python tools/profile_target.py --workload 4k16384 --frames 3 > $OUT/${TAG}_target_4k16384.log 2>&1; cat $OUT/${TAG}_target_4k16384.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_bench_8k1024_n1.json 2> $OUT/${TAG}_bench_8k1024_n1.err
echo "bench rc=$?" | tee -a $OUT/${TAG}_status.txt
head -c 400 $OUT/${TAG}_bench_8k1024_n1.json; echo
for CFG in "4 2" "8 2" "4 3" "8 3"; do
  set -- $CFG
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --emulate-world 8 --emulate-rank 1 --batch $1 --in-flight $2 \
      > $OUT/${TAG}_emulate_w8_b$1_f$2.json 2> $OUT/${TAG}_emulate_w8_b$1_f$2.err
  echo "emulate w8 batch $1 in-flight $2 rc=$?" | tee -a $OUT/${TAG}_status.txt
  head -c 160 $OUT/${TAG}_emulate_w8_b$1_f$2.json; echo
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --batch 8 --in-flight 2 > $OUT/${TAG}_bench_8k1024_n1_b8.json 2> $OUT/${TAG}_bench_8k1024_n1_b8.err
echo "bench b8 rc=$?" | tee -a $OUT/${TAG}_status.txt
head -c 160 $OUT/${TAG}_bench_8k1024_n1_b8.json; echo
