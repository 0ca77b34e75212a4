#!/bin/bash
# Runs ON the GPU box: bf16-gather tests, then PPI / large with f32 and bf16 gathered rows.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_capture.py -m gpu -q --maxfail=10 -k "bf16 or readout or oracle" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2l_pytest.log | grep -v Warn
for WL in ppi large; do for DT in f32 bf16; do
  python bench.py --steps 10 --warmup 3 --workload $WL --gather-dtype $DT --no-cpu-baseline > gpurun_out/r2l_${WL}_$DT.json 2> gpurun_out/r2l_${WL}_$DT.err; echo "$WL $DT rc=$?"; tail -2 gpurun_out/r2l_${WL}_$DT.err | grep -v bench
  python - $WL $DT <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r2l_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
print("   ", sys.argv[1], sys.argv[2], "ms", round(l["ms_per_step"],3), {x["op"].replace("b200gat_","")+":"+str(x["layer"]): round(x["ms"],3) for x in l["kernels"]}, "edge frac", round(l["edge_phase"]["frac_of_measured_hbm"],3))
PY
done; done
