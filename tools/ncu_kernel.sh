#!/bin/bash
# ncu --set full capture of one kernel from tools/ntt_only.py:  $1 = demangled-name regex, $2 = output name, $3 = skip count
set -u
PAT=$1; OUT=$2; SKIP=${3:-1}
CMD="python tools/ntt_only.py"
$CMD > gpurun_out/plain_k.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s $SKIP -c 1 -o gpurun_out/$OUT $CMD > gpurun_out/ncu_k.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_k.log
