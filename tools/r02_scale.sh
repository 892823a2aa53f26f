#!/bin/bash
# One gpurun --gpus 8 call: the strong-scaling points of the default workload launched exactly as the driver does,
# plus configs[4] (4K / 16384 spheres) at 8 GPUs.   gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_scale.sh r02'
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_scale_gpus.txt 2>&1
run() {  # N workload port
  local N=$1 WL=$2 PORT=$3
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 3 --workload $WL --no-extras --no-cpu-baseline \
      > $OUT/${TAG}_scale_${WL}_n${N}.json 2> $OUT/${TAG}_scale_${WL}_n${N}.err
  echo "scale $WL N=$N rc=$?" | tee -a $OUT/${TAG}_scale_status.txt
  head -c 240 $OUT/${TAG}_scale_${WL}_n${N}.json; echo
}
: > $OUT/${TAG}_scale_status.txt
run 8 8k1024 29621
run 8 4k16384 29622
run 4 8k1024 29623
run 2 8k1024 29624
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_scale_8k1024_n1.json 2> $OUT/${TAG}_scale_8k1024_n1.err
echo "scale 8k1024 N=1 rc=$?" | tee -a $OUT/${TAG}_scale_status.txt
