#!/bin/bash
# Round-end style check on the GPU box: gpu tests, smoke, bench (+ reference arm), launch list, full captures of the two
# top kernels (summarised on the box with tools/ncu_read.py; the first report is brought back as well).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
stamp "pytest done"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.json
stamp "bench done"
python bench.py --impl reference --steps 1 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
stamp "reference arm done"
python bench.py --bsgs --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_bsgs.json 2> gpurun_out/bench_bsgs.err; echo "bsgs rc=$?"; tail -c 400 gpurun_out/bench_bsgs.json
stamp "bsgs bench done"
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
stamp "launch list done"
n=0
for pair in "$@"; do
  PAT=${pair%%:*}; OUT=${pair##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s 40 -c 1 -f -o /tmp/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
  echo "$OUT rc=$?"
  python tools/ncu_read.py /tmp/$OUT.ncu-rep 40 > gpurun_out/$OUT.txt 2>&1
  if [ $n -eq 0 ]; then cp /tmp/$OUT.ncu-rep gpurun_out/; fi
  n=$((n+1))
  stamp "capture $OUT done"
done
