#!/bin/bash
# compute-sanitizer passes: memcheck + racecheck + initcheck-free smoke() and the streaming / queue tests.
tag=${1:-x}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/${tag}_sanitizer_${tool}_smoke.log | tail -3
  timeout 1500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_streaming_matches_oracle_as_sets or test_multi_frame_render_uploads_once or test_render_target_slack" > gpurun_out/${tag}_sanitizer_${tool}_streaming.log 2>&1
  echo "$tool streaming rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer_${tool}_streaming.log | tail -3
done
