#!/bin/bash
# Round-end style check on the GPU box: gpu tests, smoke, bench (+ reference arm), launch list, full capture of ks_digits.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KsDigitsTmem -s 30 -c 1 -o gpurun_out/prof_ksdigits $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
