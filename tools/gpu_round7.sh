#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
BRICKMAP_B200_INLINE_TESTS=1 timeout 600 python -m pytest tests -m gpu -x -q -k "traversal or adversarial or stock_world or fused or launch_frame or work_counters or caves or ragged or tiles or strips or streaming or schedules" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.log
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
X=0
BRICKMAP_B200_INLINE_TESTS=1
BRICKMAP_B200_INLINE_TESTS=2
BRICKMAP_B200_INLINE_TESTS=4
BRICKMAP_B200_INLINE_TESTS=8
BRICKMAP_B200_INLINE_TESTS=16
BRICKMAP_B200_INLINE_TESTS=1 BRICKMAP_B200_MIN_SHARE=8
BRICKMAP_B200_INLINE_TESTS=1 BRICKMAP_B200_MIN_SHARE=24
BRICKMAP_B200_INLINE_TESTS=2 BRICKMAP_B200_MIN_SHARE=8
BRICKMAP_B200_INLINE_TESTS=4 BRICKMAP_B200_MIN_SHARE=24
VARS
cat gpurun_out/${tag}_tune.log
BRICKMAP_B200_INLINE_TESTS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
