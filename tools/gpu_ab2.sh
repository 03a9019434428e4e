#!/bin/bash
# same-box A/B of library builds (round 2): bash tools/gpu_ab2.sh <tag> <lib-a> <lib-b> ... ; each run 3 times, interleaved; then parity tests on each
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_ab.log
for rep in 1 2 3; do
  for lib in "$@"; do
    BRICKMAP_B200_LIB=$PWD/brickmap_b200/$lib timeout 180 python tools/tune.py 2>&1 | tail -1 | sed "s#$PWD/brickmap_b200/##" >> gpurun_out/${tag}_ab.log
  done
done
cat gpurun_out/${tag}_ab.log
