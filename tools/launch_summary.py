"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for one denoiser pass.

    python tools/launch_summary.py gpurun_out/launches.csv [first_kernel_substring]
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
recs = []
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void dsg::<unnamed>::", "").replace("dsg::<unnamed>::", "")
    recs.append((int(x["ID"]), name, x["Grid Size"], float(x["Metric Value"].replace(",", "")) / 1000))
# one pass = from a precond_coef_kernel launch to the next node_head_kernel
starts = [i for i, r in enumerate(recs) if r[1].startswith("precond_coef")]
ends = [i for i, r in enumerate(recs) if r[1].startswith("node_head")]
s = starts[0]
e = [i for i in ends if i > s][0]
one = recs[s:e + 1]
tot = sum(r[3] for r in one)
agg = collections.OrderedDict()
for _, n, g, us in one:
    c = agg.setdefault(n, [0, 0.0])
    c[0] += 1
    c[1] += us
print(f"one denoiser pass: {len(one)} launches, {tot:.1f} us (ncu per-launch times: cold cache, serialised)")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:42s} {c:4d} {us:9.1f} us {100 * us / tot:5.1f}%")
if "-v" in sys.argv:
    for r in one:
        print(r)
