#!/bin/bash
# tools/gpu.sh [--gpus N] <timeout-seconds> '<command>' — rebuild libb200gat.so, then run the command on a B200 box via gpurun.
set -e
cd "$(dirname "$0")/.."
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
python -m atmlgraphattentionnetworks_b200.build >/dev/null
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
