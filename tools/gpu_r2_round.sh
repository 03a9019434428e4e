#!/bin/bash
# One gpurun call for a full evidence round: smoke, all GPU tests, bench (both arms), ncu launch list of the bench command, one full ncu
# capture of the frame kernel. Outputs under gpurun_out/<tag>_*.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2_round.sh <tag>' ; then: python tools/ncu_summary.py gpurun_out/<tag>_prof.ncu-rep ; python tools/ncu_traffic.py ...
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cat gpurun_out/${tag}_bench_reference.json
timeout 600 python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
cat gpurun_out/${tag}_bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:frame_kernel_q -s 20 -c 1 -f -o gpurun_out/${tag}_prof_cfg4 python tools/profile_frame.py 24 cfg4 > gpurun_out/${tag}_ncu_cfg4.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
