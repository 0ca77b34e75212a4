#!/bin/bash
# Runs ON the GPU box: the whole GPU test suite, then the default bench line; prints a digest.   usage: box_full.sh <tag> [extra bench flags]
TAG=${1:-run}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log | grep -v Warning
python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - $TAG <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
def k(r): return {x["op"].replace("b200gat_","")+":"+str(x["layer"]): round(x["ms"],3) for x in r.get("kernels",[])}
print("ppi", round(l["ms_per_step"],3), "e2e", round(l["e2e"]["ms_per_step"],3), "captured", round(l["captured"]["ms_per_step"],3), round(l["captured"]["e2e"]["ms_per_step"],3), "roof", l["roofline"]["kernel"], round(l["roofline"]["frac"],3), "edge", round(l["edge_phase"]["frac_of_measured_hbm"],3))
print("   ", k(l))
r=l["large"]; print("large", round(r["ms_per_step"],2), "e2e", round(r["e2e"]["ms_per_step"],2), "roof", r["roofline"]["kernel"], round(r["roofline"]["frac"],3), "edge", round(r["edge_phase"]["frac_of_measured_hbm"],3)); print("   ", k(r))
for n in ("cifar","cora"):
    r=l[n]; c=r["captured"]; print(n, "eager", round(r["ms_per_step"],3), "captured", round(c["ms_per_step"],3), "e2e", round(c["e2e"]["ms_per_step"],3), "launches/replay", c["gpu_launches_per_replay"]); print("   ", k(r))
c=l["cifar"]["batch512"]; print("cifar512 eager", round(c["ms_per_step"],3), "captured", round(c["captured"]["ms_per_step"],3))
for p in l["heads"]["points"]: print("heads", p["heads"], "eager", round(p["ms_per_step"],3), "captured", round(p["captured_ms_per_step"],3), "abi sum", round(p["abi_ops_ms_sum"],3), "ratio", round(p["captured_step_over_abi_ops"],2), {x["op"].replace("b200gat_",""): round(x["ms"],3) for x in p["kernels"]})
PY
