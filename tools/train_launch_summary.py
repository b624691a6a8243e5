"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list of tools/train_one_step.py: per-kernel totals of
one training iteration (every launch of the process: the library's kernels, torch's few fills / RNG kernels, memcpys are
not listed by ncu).

    python tools/train_launch_summary.py gpurun_out/train_launches.csv
"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
total, n = 0.0, 0
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void dsg::<unnamed>::", "").replace("dsg::<unnamed>::", "")
    us = float(x["Metric Value"].replace(",", "")) / 1000
    c = agg.setdefault(name[:70], [0, 0.0])
    c[0] += 1
    c[1] += us
    total += us
    n += 1
print(f"one training iteration (eager, incl. set-up kernels of the process): {n} launches, {total / 1000:.2f} ms "
      "(ncu per-launch times: cold cache, serialised)")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{name:72s} {c:4d} {us:9.1f} us {100 * us / total:5.1f}%")
