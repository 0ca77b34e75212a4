#!/bin/bash
# Runs ON a multi-GPU box: NCCL partition tests, then the row-partitioned 2.4 M-node graph at N GPUs.   usage: box_large.sh <tag> <N> [env...]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
python -m pytest tests/test_gpu_partition.py -m gpu -q -k "nccl or peer_push" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log | grep -v Warn
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload large > gpurun_out/${TAG}_large$N.json 2> gpurun_out/${TAG}_large$N.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/${TAG}_large$N.err | tail -2; tail -3 gpurun_out/${TAG}_large$N.err
python - $TAG $N <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/{sys.argv[1]}_large{sys.argv[2]}.json").read().strip().splitlines()[-1])
print("large", round(l["ms_per_step"],2), "ms", round(l["value"]/1e6,1), "M edges/s; check", l["check"]["ok"], l["check"]["loss_rel_err"], l["check"]["grad_max_rel_err"])
print("  collectives", {k: round(v,2) for k,v in l["collectives_ms_per_step"].items()}, "sum", round(sum(l["collectives_ms_per_step"].values()),2))
print("  ops", {x["op"].replace("b200gat_","")+":"+str(x["layer"]): round(x["ms"],2) for x in l["kernels"]}, "sum", round(l["abi_ops_ms_sum"],2))
PY
