#!/bin/bash
# gpurun --gpus 8: A/B of the band DMA at 8 GPUs (default: every pixel stored by the kernels; --band-dma: primary rows by copy engine
# ; measured in round 2 when the DMA was still the default and the switch was called --no-band-dma).   gpurun --gpus 8 --timeout 600 -- 'bash tools/r02_scale3.sh r02f'
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_scale_status.txt
run() {  # N workload port name extra...
  local N=$1 WL=$2 PORT=$3 NAME=$4; shift 4
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 20 --warmup 3 --workload $WL --no-extras --no-cpu-baseline "$@" \
      > $OUT/${TAG}_${NAME}.json 2> $OUT/${TAG}_${NAME}.err
  echo "$NAME rc=$?" | tee -a $OUT/${TAG}_scale_status.txt
  head -c 200 $OUT/${TAG}_${NAME}.json; echo
}
run 8 8k1024 29641 scale_8k1024_n8
run 8 8k1024 29642 scale_8k1024_n8_band_dma --band-dma
