#!/bin/bash
# ncu --set full capture of the fused inner-sum kernel of the BSGS affine layer (an HBM-bound elementwise pass)
set -u
mkdir -p gpurun_out
CMD="python bench.py --bsgs --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline"
$CMD > gpurun_out/plain_bsgs.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_bsgs.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:DyadicMacN -s 10 -c 1 -f -o /tmp/dyadic $CMD > gpurun_out/ncu_dyadic.log 2>&1
echo "rc=$?"
python tools/ncu_read.py /tmp/dyadic.ncu-rep 12 > gpurun_out/dyadic_v10.txt 2>&1
head -30 gpurun_out/dyadic_v10.txt
