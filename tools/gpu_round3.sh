#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "traversal or adversarial or stock_world or fused or launch_frame or work_counters or caves or ragged or tiles or strips" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.log
P=$PWD/brickmap_b200
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
X=0
BRICKMAP_B200_QUANTUM=64
BRICKMAP_B200_NO_FAR=1
BRICKMAP_B200_LIB=$P/libbrickmap_b200_q1024.so
BRICKMAP_B200_LIB=$P/libbrickmap_b200_q1024.so BRICKMAP_B200_QUANTUM=64
BRICKMAP_B200_LIB=$P/libbrickmap_b200_q1024.so BRICKMAP_B200_QUANTUM=96
BRICKMAP_B200_LIB=$P/libbrickmap_b200_q1024.so BRICKMAP_B200_NO_FAR=1 BRICKMAP_B200_QUANTUM=64
VARS
cat gpurun_out/${tag}_tune.log
BRICKMAP_B200_LIB=$P/libbrickmap_b200_q1024.so BRICKMAP_B200_QUANTUM=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
