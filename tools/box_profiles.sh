#!/bin/bash
# Runs ON the GPU box: the evidence captured under profiles/ — (1) launch lists (gpu__time_duration) of one train step of every
# workload, (2) one `ncu --set full` capture of every library kernel of one PPI-shaped step -> summary + per-op DRAM traffic.
# usage: box_profiles.sh <tag> <git-sha>
TAG=${1:-run}; export B200GAT_GIT_SHA=$2
RE='bwd_prep|colsum_kernel|edge_bwd_|gt_amax_kernel|bwd_finish_kernel|edge_fwd_|head_mean_kernel|amax_kernel|split_kernel|gemm_tc_kernel|gemm_simt_kernel|logits_|readout_|csr_|hub_rows|head_softmax|dropout_mask'
mkdir -p gpurun_out
for WL in ppi large cifar cora; do
  python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { echo "plain $WL failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$RE" --csv --log-file gpurun_out/${TAG}_launches_$WL.csv \
      python bench.py --profile --steps 1 --warmup 3 --workload $WL > gpurun_out/${TAG}_ncu_$WL.log 2>&1
  echo "$WL launches rc=$?"; python tools/launch_summary.py gpurun_out/${TAG}_launches_$WL.csv 12 | head -14
done
ncu --set full --clock-control none -k regex:"$RE" -s ${NCU_SKIP:-108} -c ${NCU_COUNT:-35} -f -o gpurun_out/${TAG}_prof \
    python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/${TAG}_ncufull.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_ncufull.log
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.md
python tools/ncu_traffic.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_traffic_ppi.json > gpurun_out/${TAG}_traffic.log 2>&1; tail -14 gpurun_out/${TAG}_traffic.log
rm -f gpurun_out/${TAG}_prof.ncu-rep
