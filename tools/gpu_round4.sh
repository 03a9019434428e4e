#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "traversal or adversarial or stock_world or fused or launch_frame or work_counters or caves or ragged or tiles or strips" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.log
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
X=0
BRICKMAP_B200_QUANTUM=64
BRICKMAP_B200_QUANTUM=64 BRICKMAP_B200_MIN_SHARE=12
BRICKMAP_B200_QUANTUM=64 BRICKMAP_B200_MIN_SHARE=20
BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=12
BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=16
BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=20
BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=24
BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=16
BRICKMAP_B200_QUANTUM=256 BRICKMAP_B200_MIN_SHARE=22
BRICKMAP_B200_QUANTUM=512 BRICKMAP_B200_MIN_SHARE=20
BRICKMAP_B200_QUANTUM=128 BRICKMAP_B200_MIN_SHARE=20 BRICKMAP_B200_FAR=1
VARS
cat gpurun_out/${tag}_tune.log
