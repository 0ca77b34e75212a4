#!/bin/bash
# ON the GPU box: one `ncu --set full` capture (with source) of the streaming passes of one PPI-shaped step
# (bwd_finish, bwd_prep_rows, split<ELU>): the kernels that move whole [N, D] arrays once and should run at copy speed
mkdir -p gpurun_out
python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/ns_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:'bwd_finish_kernel|bwd_prep_rows_kernel|split_kernel' -s ${SKIP:-36} -c ${COUNT:-12} -f \
    -o gpurun_out/ns_prof python bench.py --profile --steps 1 --warmup 3 --workload ppi > gpurun_out/ns_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/ns_prof.ncu-rep
