#!/bin/bash
# Run on the GPU box via gpurun. 1) the bench, 2) an ncu launch list of a short bench command, 3) one full capture of
# the dominant kernel. ncu only runs after the same command exited 0 without it.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:KsDigits -s 30 -c 2 -o gpurun_out/prof_ksdigits $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
