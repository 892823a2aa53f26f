#!/bin/bash
# One gpurun call: GPU parity tests -> executed-op capture -> bench line -> ncu launch list -> ncu --set full of the
# three default-path kernels.  Everything lands in gpurun_out/ (tag = $1, default r02).
#   gpurun --timeout 1500 -- 'bash tools/r02_measure.sh r02'
set -u
TAG=${1:-r02}
WL=${2:-8k1024}
OUT=gpurun_out
mkdir -p $OUT
OPS=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

echo "== pytest -m gpu" | tee $OUT/${TAG}_status.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_status.txt
tail -3 $OUT/${TAG}_pytest_gpu.log

echo "== profile target (no ncu)" | tee -a $OUT/${TAG}_status.txt
timeout 300 python tools/profile_target.py --workload $WL --frames 3 --json $OUT/${TAG}_units_${WL}.json > $OUT/${TAG}_target_${WL}.log 2>&1
echo "target rc=$?" | tee -a $OUT/${TAG}_status.txt
cat $OUT/${TAG}_target_${WL}.log

echo "== ncu op counters" | tee -a $OUT/${TAG}_status.txt
timeout 600 ncu --metrics $OPS --clock-control none --csv --log-file $OUT/${TAG}_ops_${WL}.csv \
    python tools/profile_target.py --workload $WL --frames 2 --json $OUT/${TAG}_ops_units_${WL}.json > $OUT/${TAG}_ops_${WL}.log 2>&1
echo "ops rc=$?" | tee -a $OUT/${TAG}_status.txt
python tools/executed_flops.py $OUT/${TAG}_ops_${WL}.csv $OUT/${TAG}_ops_units_${WL}.json $WL profiles/r02_executed_flops.json > $OUT/${TAG}_executed_${WL}.log 2>&1
cp profiles/r02_executed_flops.json $OUT/r02_executed_flops.json

echo "== bench (reference arm, then ours)" | tee -a $OUT/${TAG}_status.txt
timeout 400 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_bench_reference_arm.err
echo "bench ref rc=$?" | tee -a $OUT/${TAG}_status.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/${TAG}_bench_${WL}_n1.json 2> $OUT/${TAG}_bench_${WL}_n1.err
echo "bench rc=$?" | tee -a $OUT/${TAG}_status.txt
head -c 600 $OUT/${TAG}_bench_${WL}_n1.json; echo

echo "== ncu launch list" | tee -a $OUT/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${WL}.csv \
    python tools/profile_target.py --workload $WL --frames 3 > $OUT/${TAG}_launches_${WL}.log 2>&1
echo "launches rc=$?" | tee -a $OUT/${TAG}_status.txt

echo "== ncu --set full (one launch per kernel, frame 2)" | tee -a $OUT/${TAG}_status.txt
for K in primary_tile_kernel shade_setup_kernel shadow_sweep_kernel; do
  # the third frame's launch of each kernel (launch-skip counts matching launches only)
  # frame 0 launches 2 chunk pairs + catch-all (no hit-count hint yet), later frames one pair + the (empty) catch-all;
  # tools/ncu_summary.py summarises the longest launch of a report
  SKIP=2; COUNT=1
  [ $K = shade_setup_kernel ] && SKIP=3
  [ $K = shadow_sweep_kernel ] && SKIP=5 && COUNT=2
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip $SKIP --launch-count $COUNT \
      -o $OUT/${TAG}_ncu_${K}_${WL} -f python tools/profile_target.py --workload $WL --frames 3 > $OUT/${TAG}_ncu_${K}.log 2>&1
  echo "ncu $K rc=$?" | tee -a $OUT/${TAG}_status.txt
done
ls -la $OUT | tail -30
