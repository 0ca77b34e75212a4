#!/bin/bash
# Runs ON the GPU box: GPU tests, then PPI / large / cifar with the prep pass fused into the consumer's gX GEMM off and on.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2n_pytest.log | grep -v Warn
for V in 1 0; do
  for WL in ppi large cifar; do
    B200GAT_NO_FUSE_PREP=$V python bench.py --steps 10 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/r2n_${WL}_nofuse$V.json 2> gpurun_out/r2n_${WL}_nofuse$V.err; echo "$WL nofuse=$V rc=$?"
    python - $WL $V <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r2n_{sys.argv[1]}_nofuse{sys.argv[2]}.json").read().strip().splitlines()[-1])
c=l.get("captured") or {}
print("   ", sys.argv[1], "no_fuse", sys.argv[2], "ms", round(l["ms_per_step"],3), "captured", c.get("ms_per_step"), {x["op"].replace("b200gat_","")+":"+str(x["layer"]): round(x["ms"],3) for x in l["kernels"]})
PY
  done
done
