#!/bin/bash
# One gpurun call: parity tests, A/B throughput probes, bench, ncu captures. Outputs under gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [skip-tests]'
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -15 gpurun_out/${tag}_pytest.log
fi
: > gpurun_out/${tag}_tune.log
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
X=0
BRICKMAP_B200_NO_FAR=1
BRICKMAP_B200_QUANTUM=32
BRICKMAP_B200_QUANTUM=128
BRICKMAP_B200_SIMPLE_KERNEL=1
VARS
cat gpurun_out/${tag}_tune.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
cat gpurun_out/${tag}_bench_ours.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out | tail -12
