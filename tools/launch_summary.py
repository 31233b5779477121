"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py launches.csv [header lines...]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
agg = collections.OrderedDict()
for r in rows:
    name = r[4]
    m = re.search(r"<(?:hhe::)?(\w+)", name)
    key = m.group(1) if m and "kernel_entry" in name else name[:40]
    if "kernel_entry" in name:
        key = re.sub(r"^void kernel_entry(_c2|_c8)?<(hhe::)?", "", name).split("(T1)")[0].rstrip(">")[:60]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(v[1] for v in agg.values())
for line in sys.argv[2:]:
    print("# " + line)
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} launches={n:5d} total_us={us:12.1f} share={us / tot:.3f} avg_us={us / n:.1f}")
