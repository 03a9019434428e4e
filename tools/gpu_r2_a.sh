#!/bin/bash
# round 2, call A: lens golden from the reference build, full GPU test suite on the refactored buffer parity, baseline tune
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 300 python tests/golden/make_golden.py 256lens > gpurun_out/${tag}_golden.log 2>&1; echo "golden rc=$?"
cp gpurun_out/golden/golden_256lens.npz tests/golden/ 2>/dev/null
timeout 300 python -m pytest tests/test_oracle_golden.py -x -q -k "lens" > gpurun_out/${tag}_lens_oracle.log 2>&1; echo "lens oracle rc=$?"; tail -15 gpurun_out/${tag}_lens_oracle.log
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/${tag}_pytest.log
timeout 180 python tools/tune.py 2>&1 | tail -1 | tee gpurun_out/${tag}_tune.log
