#!/bin/bash
# ncu --set full captures of selected kernels from a short bench run, summarised on the box (reports are large):
# args = "regex:outname" ...   ->  gpurun_out/<outname>.txt (tools/ncu_read.py) ; the first report is kept as well
set -u
CMD="python bench.py --steps 1 --warmup 1 --blocks 148 --no-cpu-baseline"
$CMD > gpurun_out/plain_k.log 2>&1 || { echo "plain run failed"; exit 1; }
n=0
for pair in "$@"; do
  PAT=${pair%%:*}; OUT=${pair##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s 40 -c 1 -f -o /tmp/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
  echo "$OUT rc=$?"
  python tools/ncu_read.py /tmp/$OUT.ncu-rep 40 > gpurun_out/$OUT.txt 2>&1
  if [ $n -eq 0 ]; then cp /tmp/$OUT.ncu-rep gpurun_out/; fi
  n=$((n+1))
done
