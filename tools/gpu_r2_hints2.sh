#!/bin/bash
# Steady-state DRAM traffic (ncu single pass, --cache-control none, three consecutive launches) of library builds on cfg3 or cfg4:
# usage: gpurun -- 'bash tools/gpu_r2_hints2.sh <tag> h128:cfg3 h6:cfg4 ...'   (h<bits> = BM_SURV_HINTS build, tools/gpu_r2_hints.sh)
tag=$1; shift
mkdir -p gpurun_out
log=gpurun_out/${tag}_hints.log
: > $log
for item in "$@"; do
  v=${item%%:*}; c=${item##*:}
  if [ "$c" = tune ]; then
    echo -n "$v tune: " >> $log
    BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$v.so timeout 120 python tools/tune.py 2>&1 | tail -1 | sed 's/^{[^}]*} //' >> $log
    continue
  fi
  if [ "$c" = cfg4 ]; then frames=24; skip=20; else frames=14; skip=10; fi
  BRICKMAP_B200_LIB=$PWD/brickmap_b200/libbrickmap_b200_$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --cache-control none --clock-control none -k regex:frame_kernel_q -s $skip -c 3 --csv --log-file gpurun_out/${tag}_${v}_${c}_traffic.csv python tools/profile_frame.py $frames $c > gpurun_out/${tag}_${v}_${c}_ncu.log 2>&1
  echo "$v $c traffic (cache-control none): $(python tools/ncu_csv_rows.py gpurun_out/${tag}_${v}_${c}_traffic.csv)" >> $log
done
cat $log
