#!/bin/bash
# Same-box A/B of the survivor-set L2 hints (BM_SURV_HINTS builds libbrickmap_b200_h<bits>.so): throughput probe (tools/tune.py) and the
# DRAM traffic of three steady-state frame-kernel launches WITHOUT ncu's cache flush (--cache-control none, one pass: what the kernel
# moves when its predecessor's L2 contents are still there), then the parity tests on the most aggressive build.
# usage: gpurun --timeout 600 -- 'bash tools/gpu_r2_hints.sh <tag> h0 h1 h3 h7 h9 h5'
tag=$1; shift
mkdir -p gpurun_out
log=gpurun_out/${tag}_hints.log
: > $log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> $log
for rep in 1 2; do
  for v in "$@"; do
    echo -n "$v tune: " >> $log
    BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$v.so timeout 120 python tools/tune.py 2>&1 | tail -1 | sed 's/^{[^}]*} //' >> $log
  done
done
for v in "$@"; do
  BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$v.so timeout 180 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --cache-control none --clock-control none -k regex:frame_kernel_q -s 10 -c 3 --csv --log-file gpurun_out/${tag}_${v}_traffic.csv python tools/profile_frame.py 14 > gpurun_out/${tag}_${v}_ncu.log 2>&1
  echo "$v traffic (cache-control none): $(python tools/ncu_csv_rows.py gpurun_out/${tag}_${v}_traffic.csv)" >> $log
done
v=$1
BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$v.so timeout 180 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --cache-control all --clock-control none -k regex:frame_kernel_q -s 10 -c 3 --csv --log-file gpurun_out/${tag}_${v}_traffic_flush.csv python tools/profile_frame.py 14 > gpurun_out/${tag}_${v}_ncu_flush.log 2>&1
echo "$v traffic (cache-control all, ncu's default: L2 flushed before the launch): $(python tools/ncu_csv_rows.py gpurun_out/${tag}_${v}_traffic_flush.csv)" >> $log
cat $log
last=${PYTEST_LIB:-h7}
[ "$last" = none ] && exit 0
BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$last.so timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_$last.log 2>&1
echo "pytest $last rc=$?" | tee -a $log
tail -3 gpurun_out/${tag}_pytest_$last.log | tee -a $log
