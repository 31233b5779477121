#!/bin/bash
# ncu --set full captures of the standalone forward / inverse NTT kernels from tools/ntt_only.py, summarised on the box
set -u
mkdir -p gpurun_out
CMD="python tools/ntt_only.py"
$CMD > gpurun_out/plain_ntt.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ntt.log; exit 1; }
cat gpurun_out/plain_ntt.log | tail -3
for pair in "NttFwdCluster:nttfwd_r2a" "InvClusterBody:nttinv_r2a"; do
  PAT=${pair%%:*}; OUT=${pair##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s 2 -c 1 -f -o /tmp/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
  echo "$OUT rc=$?"
  python tools/ncu_read.py /tmp/$OUT.ncu-rep 30 > gpurun_out/$OUT.txt 2>&1
done
