#!/bin/bash
# Runs ON the GPU box: smoke(), the whole GPU suite, the default bench line (with the CPU baseline leg), a short reference-arm run
TAG=${1:-run}
python -c "import __graft_entry__ as g; g.smoke(); print(\"__SMOKE_OK__\")" 2>&1 | tail -2
bash tools/box_full.sh $TAG
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err ); echo "reference arm rc=$?"; cut -c1-300 gpurun_out/${TAG}_ref.json
