#!/bin/bash
# Runs ON the GPU box: ONLY the `ncu --set full` capture of one PPI-shaped step (all library kernels, launch order) -> summary +
# per-op DRAM traffic.   usage: box_ncu_full.sh <tag> <git-sha>      (regex without the csr_ kernels: 35 launches per step)
TAG=${1:-run}; export B200GAT_GIT_SHA=$2
RE='bwd_prep|colsum_kernel|edge_bwd_|gt_amax_kernel|bwd_finish_kernel|edge_fwd_|head_mean_kernel|amax_kernel|split_kernel|gemm_tc_kernel|gemm_simt_kernel|logits_'
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none -k regex:"$RE" -s ${NCU_SKIP:-105} -c ${NCU_COUNT:-35} -f -o gpurun_out/${TAG}_prof \
    python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/${TAG}_ncufull.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_ncufull.log
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.md
python tools/ncu_traffic.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_traffic_ppi.json > gpurun_out/${TAG}_traffic.log 2>&1; tail -14 gpurun_out/${TAG}_traffic.log
rm -f gpurun_out/${TAG}_prof.ncu-rep
