#!/bin/bash
# gpurun with retries on "transient" (pod busy): tools/grun.sh <logfile> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 40); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "exit code 3" "$log"; then sleep 45; continue; fi
  break
done
echo "done attempt=$attempt" >> "$log"
