#!/bin/bash
# Runs ON the GPU box: ncu launch lists (gpu__time_duration) of ONE eager train step of the launch-bound workloads.
mkdir -p gpurun_out
for WL in cifar cora; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_$WL.csv \
      python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/r2f_ncu_$WL.log 2>&1
  echo "$WL rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_heads1.csv \
      python bench.py --profile --steps 1 --warmup 3 --workload heads --heads 1 > gpurun_out/r2f_ncu_heads1.log 2>&1
echo "heads1 rc=$?"
