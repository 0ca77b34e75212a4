#!/bin/bash
# Runs ON the GPU box: GPU tests, then the PPI / large bench with the gW side-stream overlap off and on.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log | grep -v Warn
for V in 0 1; do
  for WL in ppi large; do
    B200GAT_OVERLAP_GW=$V python bench.py --steps 20 --warmup 5 --workload $WL --no-cpu-baseline > gpurun_out/r2j_${WL}_ov$V.json 2> gpurun_out/r2j_${WL}_ov$V.err; echo "$WL overlap=$V rc=$?"
    python - $WL $V <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r2j_{sys.argv[1]}_ov{sys.argv[2]}.json").read().strip().splitlines()[-1])
c=l.get("captured") or {}
print("   ", sys.argv[1], "overlap", sys.argv[2], "ms", round(l["ms_per_step"],3), "e2e", round(l["e2e"]["ms_per_step"],3), "captured", c.get("ms_per_step"))
PY
  done
done
