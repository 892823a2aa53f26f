#!/bin/bash
# One gpurun call (1 GPU): the other BASELINE.json configs, the per-rank emulation of 4- and 8-rank jobs, and the
# ncu --set full capture of shade_setup_kernel.   gpurun --timeout 1200 -- 'bash tools/r02_configs.sh r02'
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
echo "== configs" | tee $OUT/${TAG}_configs_status.txt
for WL in vga64 1080p64 4k1024 4k16384; do
  timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $WL --no-extras > $OUT/${TAG}_bench_${WL}_n1.json 2> $OUT/${TAG}_bench_${WL}_n1.err
  echo "bench $WL rc=$?" | tee -a $OUT/${TAG}_configs_status.txt
  head -c 300 $OUT/${TAG}_bench_${WL}_n1.json; echo
done
echo "== emulated ranks" | tee -a $OUT/${TAG}_configs_status.txt
for EW in 2 4 8; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --emulate-world $EW --emulate-rank 1 > $OUT/${TAG}_emulate_8k1024_w${EW}.json 2> $OUT/${TAG}_emulate_8k1024_w${EW}.err
  echo "emulate $EW rc=$?" | tee -a $OUT/${TAG}_configs_status.txt
  head -c 300 $OUT/${TAG}_emulate_8k1024_w${EW}.json; echo
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline --workload 4k16384 --emulate-world 8 --emulate-rank 1 > $OUT/${TAG}_emulate_4k16384_w8.json 2> $OUT/${TAG}_emulate_4k16384_w8.err
echo "emulate 4k16384 w8 rc=$?" | tee -a $OUT/${TAG}_configs_status.txt
echo "== ncu shade_setup" | tee -a $OUT/${TAG}_configs_status.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:shade_setup_kernel --launch-skip 3 --launch-count 1 \
    -o $OUT/${TAG}_ncu_shade_setup_kernel_8k1024 -f python tools/profile_target.py --workload 8k1024 --frames 3 > $OUT/${TAG}_ncu_shade_setup_kernel.log 2>&1
echo "ncu shade_setup rc=$?" | tee -a $OUT/${TAG}_configs_status.txt
timeout 300 python tools/profile_target.py --workload 4k16384 --frames 3 > $OUT/${TAG}_target_4k16384.log 2>&1
cat $OUT/${TAG}_target_4k16384.log
