#!/bin/bash
# Runs ON the GPU box: the whole GPU suite, the default bench line, a short reference-arm run, then the ncu --set full capture
# of one PPI-shaped step (summary + per-op DRAM traffic).  usage: box_final.sh <tag> <git-sha>
TAG=${1:-run}; export B200GAT_GIT_SHA=$2
python -c "import __graft_entry__ as g; g.smoke(); print(\"__SMOKE_OK__\")" 2>&1 | tail -3
bash tools/box_full.sh $TAG
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err ); echo "reference arm rc=$?"; cut -c1-400 gpurun_out/${TAG}_ref.json
RE='bwd_prep|colsum_kernel|edge_bwd_|gt_amax_kernel|bwd_finish_kernel|edge_fwd_|head_mean_kernel|amax_kernel|split_kernel|gemm_tc_kernel|gemm_simt_kernel|logits_'
ncu --set full --clock-control none -k regex:"$RE" -s ${NCU_SKIP:-105} -c ${NCU_COUNT:-35} -f -o gpurun_out/${TAG}_prof \
    python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/${TAG}_ncufull.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/${TAG}_ncufull.log
python tools/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.md
python tools/ncu_traffic.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_traffic_ppi.json > gpurun_out/${TAG}_traffic.log 2>&1; tail -14 gpurun_out/${TAG}_traffic.log
rm -f gpurun_out/${TAG}_prof.ncu-rep
