#!/bin/bash
# Runs ON the GPU box: one `ncu --set full` capture of the hot kernels of the LAST timed step of `bench.py --profile`.
# usage: box_ncu_full.sh <tag> <skip> <count> [regex]
TAG=${1:-run}; SKIP=${2:-45}; COUNT=${3:-15}; RE=${4:-gemm_tc|edge_fwd_kernel|edge_bwd_kernel}
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"$RE" -s $SKIP -c $COUNT \
    -f -o gpurun_out/${TAG}_prof python bench.py --profile --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu2.log
# the report embeds the library's cubin and exceeds gpurun's 64 MiB return limit: summarise it here, keep the summaries
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.md
python tools/ncu_traffic.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_traffic.log 2>&1
rm -f gpurun_out/${TAG}_prof.ncu-rep
