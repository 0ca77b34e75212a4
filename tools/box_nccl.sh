#!/bin/bash
# Runs ON an 8-GPU box: the row-partitioned 2.4 M-node graph under a few NCCL settings (all-gather bandwidth is what bounds it)
mkdir -p gpurun_out
run() { TAG=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --workload large > gpurun_out/r2v_$TAG.json 2> gpurun_out/r2v_$TAG.err
  python - $TAG <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r2v_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(l["ms_per_step"],2), "ms; coll", {k: round(v,2) for k,v in l["collectives_ms_per_step"].items()}, "check", l["check"]["ok"])
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL
grep -i "nvls\|channels\|Connected\|AllGather.*algo\|algo" gpurun_out/r2v_default.err | grep -v "^\[bench" | sort | uniq -c | sort -rn | head -12 | cut -c1-220
run minch32 NCCL_MIN_NCHANNELS=32
run nvls0 NCCL_NVLS_ENABLE=0
run ctas NCCL_MIN_CTAS=32 NCCL_MAX_CTAS=64
