#!/bin/bash
# The evidence round of the final tree in ONE gpurun call: smoke, all GPU tests, both bench arms, the ncu launch list of the bench command,
# full ncu captures of the frame kernel (cfg3, cfg4) plus single-pass steady-state traffic captures, the traffic stamp written ON the box
# and a second bench line that carries it, then every BASELINE config as a bench line. Outputs under gpurun_out/<tag>_*.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_r2_final.sh <tag>' ; then copy gpurun_out/<tag>_frame_kernel_traffic.json to profiles/frame_kernel_traffic.json
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
timeout 400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
head -c 300 gpurun_out/${tag}_bench_reference.json; echo
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -f -o gpurun_out/${tag}_prof python tools/profile_frame.py > gpurun_out/${tag}_ncu.log 2>&1
timeout 200 ncu --metrics $M --cache-control none --clock-control none -k regex:frame_kernel_q -s 10 -c 3 --csv --log-file gpurun_out/${tag}_steady_cfg3.csv python tools/profile_frame.py 14 > gpurun_out/${tag}_steady_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:frame_kernel_q -s 20 -c 1 -f -o gpurun_out/${tag}_prof_cfg4 python tools/profile_frame.py 24 cfg4 > gpurun_out/${tag}_ncu_cfg4.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:frame_kernel_q -s 20 -c 3 --csv --log-file gpurun_out/${tag}_steady_cfg4.csv python tools/profile_frame.py 24 cfg4 > gpurun_out/${tag}_steady_cfg4.log 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}_prof.ncu-rep cfg3 - gpurun_out/${tag}_steady_cfg3.csv > /dev/null 2> gpurun_out/${tag}_traffic.err
python tools/ncu_traffic.py gpurun_out/${tag}_prof_cfg4.ncu-rep cfg4 - gpurun_out/${tag}_steady_cfg4.csv > /dev/null 2>> gpurun_out/${tag}_traffic.err
cp profiles/frame_kernel_traffic.json gpurun_out/${tag}_frame_kernel_traffic.json
timeout 300 python bench.py > gpurun_out/${tag}_bench_ours.json 2> gpurun_out/${tag}_bench_ours.err
cat gpurun_out/${tag}_bench_ours.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"
# every other BASELINE config (reference first)
for c in cfg2 tour cfg5; do
  timeout 400 python bench.py --config $c --impl reference > gpurun_out/${tag}_bench_${c}_reference.json 2> gpurun_out/${tag}_bench_${c}_reference.err; echo "$c reference rc=$?"
  timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/${tag}_bench_${c}_ours.json 2> gpurun_out/${tag}_bench_${c}_ours.err; echo "$c ours rc=$?"
  head -c 250 gpurun_out/${tag}_bench_${c}_ours.json; echo
done
timeout 300 python bench.py --config cfg3 --spp 1 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg3_1spp_ours.json 2> gpurun_out/${tag}_bench_cfg3_1spp_ours.err
timeout 300 python bench.py --config cfg3 --spp 1 --impl reference > gpurun_out/${tag}_bench_cfg3_1spp_reference.json 2> gpurun_out/${tag}_bench_cfg3_1spp_reference.err
timeout 600 python bench.py --config cfg4 > gpurun_out/${tag}_bench_cfg4_ours.json 2> gpurun_out/${tag}_bench_cfg4_ours.err; echo "cfg4 rc=$?"
head -c 250 gpurun_out/${tag}_bench_cfg4_ours.json; echo
