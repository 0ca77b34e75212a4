#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_geometry.py -m gpu -q --maxfail=10 -k "stack or gatnet" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log | grep -v Warn
for V in 0 1; do
  for WL in ppi large; do
    B200GAT_NO_FUSE_PREP=$V python bench.py --steps 10 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/r2o_${WL}_nofuse$V.json 2> gpurun_out/r2o_${WL}_nofuse$V.err; echo "$WL nofuse=$V rc=$?"
    python - $WL $V <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r2o_{sys.argv[1]}_nofuse{sys.argv[2]}.json").read().strip().splitlines()[-1])
print("   ", sys.argv[1], "no_fuse", sys.argv[2], "ms", round(l["ms_per_step"],3), {x["op"].replace("b200gat_","")+":"+str(x["layer"]): round(x["ms"],3) for x in l["kernels"] if "bwd" in x["op"]})
PY
  done
done
