#!/bin/bash
# Runs ON a multi-GPU box: the default bench line (all workloads) at N GPUs through torchrun, as the driver launches it.
N=${1:-8}; TAG=${2:-scale}
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err ); echo "bench rc=$?"
grep "\[bench\]" gpurun_out/${TAG}_bench$N.err | tail -8; tail -4 gpurun_out/${TAG}_bench$N.err
python - $N $TAG <<'PY'
import json, sys
N, TAG = sys.argv[1], sys.argv[2]
l=json.loads(open(f'gpurun_out/{TAG}_bench{N}.json').read().strip().splitlines()[-1])
print('ppi', round(l['ms_per_step'],3), round(l['value']/1e6,1), 'M edges/s check', l.get('check',{}).get('ok'))
r=l['large']; print('large', round(r['ms_per_step'],2), round(r['value']/1e6,1), 'check', r['check']['ok'], 'coll', {k: round(v,2) for k,v in r['collectives_ms_per_step'].items()})
r=l['bf16_gather']['large']; print('large bf16', round(r['ms_per_step'],2), round(r['value']/1e6,1), 'check', r['check']['ok'], 'coll', {k: round(v,2) for k,v in r['collectives_ms_per_step'].items()})
for k,r in l['cifar'].items(): print('cifar', k, round(r['ms_per_step'],3), round(r['value']/1e6,1), 'check', r['check']['ok'])
PY
