#!/bin/bash
# every BASELINE config as a bench line on one GPU, both arms (reference first): gpurun --timeout 3000 -- 'bash tools/gpu_bench_all.sh <tag>'
tag=${1:-x}
mkdir -p gpurun_out
for c in cfg3 cfg2 tour cfg5; do
  timeout 900 python bench.py --config $c --impl reference > gpurun_out/${tag}_bench_${c}_reference.json 2> gpurun_out/${tag}_bench_${c}_reference.err; echo "$c reference rc=$?"
  head -c 600 gpurun_out/${tag}_bench_${c}_reference.json; echo
  timeout 900 python bench.py --config $c > gpurun_out/${tag}_bench_${c}_ours.json 2> gpurun_out/${tag}_bench_${c}_ours.err; echo "$c ours rc=$?"; tail -2 gpurun_out/${tag}_bench_${c}_ours.err
  head -c 600 gpurun_out/${tag}_bench_${c}_ours.json; echo
done
timeout 900 python bench.py --config cfg3 --spp 1 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg3_1spp_ours.json 2> gpurun_out/${tag}_bench_cfg3_1spp_ours.err; head -c 400 gpurun_out/${tag}_bench_cfg3_1spp_ours.json; echo
timeout 900 python bench.py --config cfg3 --spp 1 --impl reference > gpurun_out/${tag}_bench_cfg3_1spp_reference.json 2> gpurun_out/${tag}_bench_cfg3_1spp_reference.err; head -c 400 gpurun_out/${tag}_bench_cfg3_1spp_reference.json; echo
timeout 1500 python bench.py --config cfg4 > gpurun_out/${tag}_bench_cfg4_ours.json 2> gpurun_out/${tag}_bench_cfg4_ours.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/${tag}_bench_cfg4_ours.err; cat gpurun_out/${tag}_bench_cfg4_ours.json
