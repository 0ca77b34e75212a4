#!/bin/bash
# Runs ON the GPU box: parity tests, then short bench lines of the named workloads (no CPU leg).
# usage: box_quick.sh <tag> [workload ...]
TAG=${1:-q}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_test.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_test.log
for WL in "$@"; do
  python bench.py --workload $WL --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${TAG}_${WL}.json 2> gpurun_out/${TAG}_${WL}.err
  echo "bench $WL rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${WL}.json").read().strip().splitlines()[-1])
    print("$WL", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e6, 1), "Medges/s e2e", round(d["e2e"]["ms_per_step"], 3), "edge_phase", d.get("edge_phase", {}).get("frac_of_measured_hbm"))
    for k in d["kernels"]:
        print("  ", k["op"], k["layer"], k["geom"], round(k["ms"], 4), "ms", round(k["frac_hbm"], 3))
except Exception as e:
    print("parse failed", e)
P
done
