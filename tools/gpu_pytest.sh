#!/bin/bash
# usage: gpurun -- 'bash tools/gpu_pytest.sh <tag> [pytest args]'
tag=${1:-t}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short --durations=8 "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/${tag}_pytest.log | cut -c1-250
