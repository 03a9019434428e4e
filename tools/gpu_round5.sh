#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
BRICKMAP_B200_MODE=2 timeout 600 python -m pytest tests -m gpu -x -q -k "traversal or adversarial or stock_world or fused or launch_frame or work_counters or caves or ragged or tiles or strips or streaming" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.log
P=$PWD/brickmap_b200
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
BRICKMAP_B200_QUANTUM=64
BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=64
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=16
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=22
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16 BRICKMAP_B200_LIB=$P/libbrickmap_b200_c4.so
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16 BRICKMAP_B200_LIB=$P/libbrickmap_b200_c16.so
BRICKMAP_B200_MODE=0 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16 BRICKMAP_B200_LIB=$P/libbrickmap_b200_c16.so
VARS
sed "s#$P/##" gpurun_out/${tag}_tune.log
BRICKMAP_B200_MODE=2 BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
