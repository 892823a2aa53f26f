#!/bin/bash
# gpurun --gpus 8: the 8-GPU points after the dark-light / sky-shortcut kernels, with one diagnostic run that keeps every
# rank's rows in its own memory (isolates the NVLink stores).   gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_scale2.sh r02d'
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_scale_status.txt
run() {  # N workload port name extra...
  local N=$1 WL=$2 PORT=$3 NAME=$4; shift 4
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 3 --workload $WL --no-extras --no-cpu-baseline "$@" \
      > $OUT/${TAG}_${NAME}.json 2> $OUT/${TAG}_${NAME}.err
  echo "$NAME rc=$?" | tee -a $OUT/${TAG}_scale_status.txt
  head -c 200 $OUT/${TAG}_${NAME}.json; echo
}
run 8 8k1024 29631 scale_8k1024_n8
run 8 8k1024 29632 diag_local_frames_8k1024_n8 --diag-local-frames
run 8 4k16384 29633 scale_4k16384_n8
run 8 8k1024 29634 scale_8k1024_n8_steps80 --steps 80
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_scale_8k1024_n1.json 2> $OUT/${TAG}_scale_8k1024_n1.err
echo "scale_8k1024_n1 rc=$?" | tee -a $OUT/${TAG}_scale_status.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi_after.txt 2>&1
