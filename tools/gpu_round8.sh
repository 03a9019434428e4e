#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python tools/bench_views.py --frames 32 --reference > gpurun_out/${tag}_camera_tour.txt 2>&1; tail -14 gpurun_out/${tag}_camera_tour.txt
timeout 600 python tools/run_caves.py 8192 64 > gpurun_out/${tag}_caves_8192.txt 2>&1; tail -8 gpurun_out/${tag}_caves_8192.txt
