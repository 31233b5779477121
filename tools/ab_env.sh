#!/bin/bash
# A/B of engine variants selected by environment switches: tools/ab_env.sh "" "HHE_NO_PREFETCH=1" "HHE_STRICT_CLUSTER=1" ...
# Each argument is a (possibly empty) list of VAR=value settings; one short bench run per setting, summary lines at the end.
set -u
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || { echo "cfg '$cfg' failed"; tail -5 gpurun_out/ab_$i.err; }
  python - "$cfg" gpurun_out/ab_$i.json <<'PY'
import json, sys
cfg, path = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(path))
except Exception as e:
    print(f"[{cfg}] no result: {e}"); sys.exit(0)
km = d["roofline"]["kernel_ms"]
top = {k: round(v / d["steps"], 1) for k, v in list(km.items())[:8]}
print(f"[{cfg or 'default'}] value {d['value']:.2f} e2e {d['e2e']['value']:.2f} ms/step {d['ms_per_step']:.1f} ntt {d['ntt']['fwd']['GBps']:.0f}/{d['ntt']['inv']['GBps']:.0f} GB/s  per-step kernel ms {top} sm_mhz {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
  i=$((i+1))
done
