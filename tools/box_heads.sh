#!/bin/bash
# Runs ON the GPU box: heads sweep (BASELINE configs[3]) — per-op times of one 50 -> H x 64 layer on the PPI-shaped batch.
TAG=${1:-hs}
mkdir -p gpurun_out
for H in 1 2 4 8 16; do
  python bench.py --workload heads --heads $H --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_h$H.json 2> gpurun_out/${TAG}_h$H.err
  python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_h$H.json").read().strip().splitlines()[-1])
print("H=$H", round(d["ms_per_step"], 3), "ms", " ".join(f'{k["op"][8:]}={k["ms"]*1e3:.0f}us' for k in d["kernels"]))
P
done
