#!/bin/bash
# N-GPU bench lines (run under gpurun --gpus N): bash tools/gpu_scale.sh <tag> <N> [configs...]   e.g. "cfg3 cfg3:1 cfg5"  (cfgX:S = --spp S)
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
for item in "$@"; do
  c=${item%%:*}; spp=""; name=$c
  if [[ "$item" == *:* ]]; then spp="--spp ${item##*:}"; name="${c}_${item##*:}spp"; fi
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --config $c $spp --gpus 1 --no-cpu-baseline > gpurun_out/${tag}_${name}_${n}gpu.json 2> gpurun_out/${tag}_${name}_${n}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --config $c $spp --gpus $n --no-cpu-baseline > gpurun_out/${tag}_${name}_${n}gpu.json 2> gpurun_out/${tag}_${name}_${n}gpu.err
  fi
  echo "$name x$n rc=$?"; tail -2 gpurun_out/${tag}_${name}_${n}gpu.err | cut -c1-300; head -c 700 gpurun_out/${tag}_${name}_${n}gpu.json; echo
done
