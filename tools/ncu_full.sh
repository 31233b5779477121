#!/bin/bash
# one ncu --set full capture of the kernel matching $1 (demangled-name regex), short bench command
set -u
PAT=${1:-KsDigits}; OUT=${2:-prof}; SKIP=${3:-30}
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s $SKIP -c 2 -o gpurun_out/$OUT $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | grep ncu-rep
