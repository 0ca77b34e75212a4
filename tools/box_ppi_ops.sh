#!/bin/bash
# one PPI bench line, reduced to: ms/step, the parity check record, per-op times
python bench.py --steps ${STEPS:-10} --warmup 3 --workload ${WL:-ppi} --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
chk=[v for k,v in l.items() if k=='check'] or [l.get('config',{}).get('check')]
print(round(l['ms_per_step'],3), 'check', chk, {x['op'].replace('b200gat_','')+':'+str(x['layer']): round(x['ms'],3) for x in l.get('kernels',[])})"
