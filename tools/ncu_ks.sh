#!/bin/bash
# ncu --set full capture of one ks_digits launch (148 items) from a short bench run + a second pass with the counters the bench's
# roofline object quotes (FP64 warp instructions, L2->SM bytes, DRAM bytes). usage: tools/ncu_ks.sh <tag>
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --blocks 148 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/plain_ks.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ks.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KsDigitsTmem -s 40 -c 1 -f -o /tmp/ks_$TAG $CMD > gpurun_out/ncu_ks_$TAG.log 2>&1
echo "full capture rc=$?"
python tools/ncu_read.py /tmp/ks_$TAG.ncu-rep 40 > gpurun_out/ksdigits_${TAG}_ncu.txt 2>&1
ncu --clock-control none --kernel-name-base demangled -k regex:KsDigitsTmem -s 40 -c 1 --csv \
    --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    $CMD > gpurun_out/ksdigits_${TAG}_counters.csv 2> gpurun_out/ncu_ks2_$TAG.log
echo "counter pass rc=$?"
grep -E "KsDigits" gpurun_out/ksdigits_${TAG}_counters.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
