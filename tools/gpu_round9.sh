#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
: > gpurun_out/${tag}_tune.log
while read -r v; do
  env $v timeout 180 python tools/tune.py 2>&1 | tail -1 >> gpurun_out/${tag}_tune.log
done <<VARS
X=0
X=1
BRICKMAP_B200_SIMPLE_KERNEL=1
VARS
cat gpurun_out/${tag}_tune.log
