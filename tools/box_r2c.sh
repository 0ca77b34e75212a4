#!/bin/bash
# Runs ON a multi-GPU box: NCCL partition tests, then the default bench line at N GPUs through torchrun.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_partition.py -m gpu -q -k "nccl or peer_push" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2c_bench$N.json 2> gpurun_out/r2c_bench$N.err ); echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2c_bench$N.err | tail -12; tail -5 gpurun_out/r2c_bench$N.err
python - $N <<'PY'
import json, sys
N=sys.argv[1]
try:
    l=json.loads(open(f'gpurun_out/r2c_bench{N}.json').read().strip().splitlines()[-1])
    print('ppi', l['ms_per_step'], l['value'], 'check', l.get('check'))
    r=l['large']; print('large', r['ms_per_step'], r['value'], 'check', r.get('check'), 'coll', r.get('collectives_ms_per_step'))
    for k,r in l['cifar'].items(): print('cifar', k, r['ms_per_step'], r['value'], 'check', r.get('check'))
except Exception as e:
    print('parse failed', e)
PY
