#!/bin/bash
# compute-sanitizer racecheck only (the shipped tree's discard / cache-hint paths): smoke() and the multi-frame bm_render tests
tag=${1:-x}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 100 $CS --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/${tag}_sanitizer_racecheck_smoke.log | tail -3
timeout 100 $CS --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_streaming_matches_oracle_as_sets or test_multi_frame_render_uploads_once or test_render_target_slack" > gpurun_out/${tag}_sanitizer_racecheck_streaming.log 2>&1
echo "racecheck streaming rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer_racecheck_streaming.log | tail -3
