#!/bin/bash
# Round-2 evidence run on one B200: default bench line (with the secondary configs), reference arm, ncu launch list of a short bench
# command, ncu --set full captures of the top kernels (ks_digits, rot_tail, forward / inverse NTT, the cluster-8 small-request kernel).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench.err
stamp "bench done"
python bench.py --impl reference --steps 1 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
stamp "reference arm done"
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1
stamp "launch list done"
for pair in "KsDigitsTmem:r2_ksdigits" "RotTail:r2_rottail"; do
  PAT=${pair%%:*}; OUT=${pair##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s 40 -c 1 -f -o /tmp/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
  echo "$OUT rc=$?"
  python tools/ncu_read.py /tmp/$OUT.ncu-rep 40 > gpurun_out/${OUT}_ncu.txt 2>&1
  stamp "capture $OUT done"
done
cp /tmp/r2_ksdigits.ncu-rep gpurun_out/ 2>/dev/null
NTT="python tools/ntt_only.py"
$NTT > gpurun_out/plain_ntt.log 2>&1
for pair in "NttFwdCluster:r2_nttfwd" "InvClusterBody:r2_nttinv"; do
  PAT=${pair%%:*}; OUT=${pair##*:}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$PAT -s 2 -c 1 -f -o /tmp/$OUT $NTT > gpurun_out/ncu_$OUT.log 2>&1
  echo "$OUT rc=$?"
  python tools/ncu_read.py /tmp/$OUT.ncu-rep 30 > gpurun_out/${OUT}_ncu.txt 2>&1
done
stamp "ntt captures done"
ONE="python bench.py --steps 1 --warmup 3 --blocks 1 --no-cpu-baseline --no-configs"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KsDigitsSplit -s 100 -c 1 -f -o /tmp/r2_kssplit $ONE > gpurun_out/ncu_r2_kssplit.log 2>&1
python tools/ncu_read.py /tmp/r2_kssplit.ncu-rep 25 > gpurun_out/r2_kssplit_ncu.txt 2>&1
stamp "split capture done"
