#!/bin/bash
# Runs ON the GPU box: `ncu --set full` capture of the edge kernels of the last step of `bench.py --profile`.
# usage: box_ncu_edge.sh <tag> <skip> <count> <regex> [workload]
TAG=${1:-run}; SKIP=${2:-18}; COUNT=${3:-6}; RE=${4:-edge_fwd|edge_bwd}; WL=${5:-ppi}
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $COUNT \
    -f -o gpurun_out/${TAG}_prof python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu2.log
