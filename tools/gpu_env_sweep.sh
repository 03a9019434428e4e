#!/bin/bash
# env-switch sweep of the scheduling parameters with tools/tune.py: bash tools/gpu_env_sweep.sh <tag> "VAR=a VAR2=b" "VAR=c" ...
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_sweep.log
for rep in 1 2; do
  for v in "$@"; do
    env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_sweep.log
  done
done
cat gpurun_out/${tag}_sweep.log
