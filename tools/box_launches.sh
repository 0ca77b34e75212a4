#!/bin/bash
# Runs ON the GPU box: per-launch device times of selected kernels of one timed step of `bench.py --profile`.
# usage: box_launches.sh <tag> <regex> [workload]
TAG=${1:-run}; RE=${2:-.}; WL=${3:-ppi}
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$RE" --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "rc=$?"; cat gpurun_out/${TAG}_plain.log | tail -1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 20
