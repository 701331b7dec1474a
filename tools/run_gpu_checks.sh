#!/bin/bash
# host bandwidth diagnosis + e2e phases at several staging thread counts
mkdir -p gpurun_out
timeout 300 python tools/host_bw.py > gpurun_out/host_bw.log 2>&1
cat gpurun_out/host_bw.log
rm -f gpurun_out/copy_ab.log
for t in 1 2 4 8; do
  echo "== J2P_COPY_THREADS=$t" >> gpurun_out/copy_ab.log
  J2P_COPY_THREADS=$t timeout 200 python tools/e2e_trace.py 2>&1 | grep -v "^$" | tail -7 >> gpurun_out/copy_ab.log
done
cat gpurun_out/copy_ab.log
