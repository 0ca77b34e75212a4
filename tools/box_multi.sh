#!/bin/bash
# Runs ON the GPU box (gpurun --gpus N): the driver's torchrun launch of bench.py for the named workloads.
# usage: box_multi.sh <tag> <N> wl...
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
for WL in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $WL --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_${WL}$N.json 2> gpurun_out/${TAG}_${WL}$N.err
  echo "rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${WL}$N.json").read().strip().splitlines()[-1])
    print("$WL N=$N", d["config"]["parallelism"][:30], round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e6, 1), "Medges/s e2e", round(d["e2e"]["ms_per_step"], 2), "scaling", d["scaling"])
    print("  ", " ".join(f'{k["op"][8:]}{k["layer"]}={k["ms"]:.2f}' for k in d["kernels"]))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/${TAG}_${WL}$N.err").read()[-1500:])
P
done
