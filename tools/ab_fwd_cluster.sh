set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_full3.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fwdcluster.json 2> gpurun_out/bench_fwdcluster.err; tail -c 900 gpurun_out/bench_fwdcluster.json
HHE_NO_FWD_CLUSTER=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nofwdcluster.json 2>/dev/null; tail -c 900 gpurun_out/bench_nofwdcluster.json
python bench.py --bsgs --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bsgs2.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/bench_bsgs2.json'));print('bsgs',d['value'],d['e2e']['value'])"
