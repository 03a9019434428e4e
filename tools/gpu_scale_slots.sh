#!/bin/bash
# frame size sweep for the small-tile case (1080p x 16 spp over N GPUs): bash tools/gpu_scale_slots.sh <tag> <N> <slots...>
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
for s in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --config cfg3 --slots $s --gpus $n --no-cpu-baseline > gpurun_out/${tag}_cfg3_slots${s}_${n}gpu.json 2> gpurun_out/${tag}_cfg3_slots${s}_${n}gpu.err
  echo "slots $s x$n rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/${tag}_cfg3_slots${s}_${n}gpu.json'));print(d['value'],d['ms_per_step'],d['config']['frames_per_step'],d['config']['rays_per_step'],d['roofline']['kernel_share_of_step'])"
done
