#!/bin/bash
# Runs ON the GPU box (via tools/gpu.sh): parity tests, smoke, the default bench line, an ncu launch list of one step and
# one `--set full` capture of every library kernel of one step.  Outputs land in gpurun_out/<tag>_*.
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_test.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_test.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --profile --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
RE='amax_kernel|split_kernel|gemm_|logits_kernel|edge_|head_mean|bwd_|colsum|gt_amax'
N=$(python tools/launch_count.py gpurun_out/${TAG}_launches.csv "$RE" 5)
echo "library kernels per step: $N"
bash tools/box_ncu_full.sh ${TAG} $((3 * N)) $N "$RE"
