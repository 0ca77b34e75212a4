#!/bin/bash
# Runs ON the GPU box: GPU test suite, then the default bench line (all workloads) and the reference arm on a small sample.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err ); echo "bench rc=$?"
tail -12 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
    print('ppi', l['ms_per_step'], l['value'], 'e2e', l['e2e']['ms_per_step'], 'roof', l['roofline']['kernel'], l['roofline']['frac'])
    for k in ('large','cifar','cora'):
        r=l[k]; print(k, r['ms_per_step'], r['value'], 'e2e', r.get('e2e',{}).get('ms_per_step'), 'roof', (r.get('roofline') or {}).get('kernel'), (r.get('roofline') or {}).get('frac'), 'edge', (r.get('edge_phase') or {}).get('frac_of_measured_hbm'), 'cpu', (r.get('cpu_baseline') or {}).get('value'), 'wall', r.get('wall_s'))
    print('cifar512', l['cifar']['batch512']['ms_per_step'])
    for p in l['heads']['points']: print('heads', p['heads'], p['ms_per_step'], p['abi_ops_ms_sum'], p['step_over_abi_ops'])
except Exception as e:
    print('parse failed', e)
PY
