#!/bin/bash
# Runs ON the GPU box: A/B of an environment switch on several workloads.  usage: box_ab.sh <tag> "<ENV=VAL>" wl...
TAG=$1; ENVV=$2; shift 2
mkdir -p gpurun_out
for WL in "$@"; do
  for MODE in A B; do
    if [ $MODE = B ]; then export $ENVV; fi
    python bench.py --workload $WL --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_${WL}_$MODE.json 2> gpurun_out/${TAG}_${WL}_$MODE.err
    python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_${WL}_$MODE.json").read().strip().splitlines()[-1])
print("$WL $MODE", round(d["ms_per_step"], 3), "ms", " ".join(f'{k["op"][8:]}{k["layer"]}={k["ms"]*1e3:.0f}' for k in d["kernels"]))
P
  done
  unset ${ENVV%%=*}
done
