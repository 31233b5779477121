#!/bin/bash
# quick GPU check after a kernel change: gpu tests, then the bench (kernel-time table + NTT GB/s)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_quick.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
print("kernel_ms", d["roofline"]["kernel_ms"])
print("ntt", {k: round(v["GBps"], 1) for k, v in d["ntt"].items() if isinstance(v, dict)})
print("ks_digits avg ms", d["roofline"]["avg_launch_ms"], "fp64 frac", d["roofline"]["frac"], "clocks", d["clocks"])
PY
