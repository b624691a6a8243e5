"""Per-kernel totals of one denoiser pass from an ncu launch list (gpu__time_duration, dram bytes, tensor-pipe activity).

    python tools/ncu_pass_summary.py profiles/r1_launches_pass_ncu.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = L.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(L)
starts = [i for i in ids if "sinusoid_kernel" in L[i]["name"]]
seq = [i for i in ids if i >= starts[-1] and "at::" not in L[i]["name"]]


def short(n):
    m = re.search(r"(\w+_kernel(<[^>]*>)?)", n)
    return m.group(1) if m else n[:40]


agg, tot = collections.OrderedDict(), 0.0
for i in seq:
    e = L[i]
    a = agg.setdefault(short(e["name"]), [0, 0.0, 0.0, 0.0, 0.0])
    t = e["gpu__time_duration.sum"]
    a[0] += 1; a[1] += t; a[2] += e["dram__bytes_read.sum"]; a[3] += e["dram__bytes_write.sum"]
    a[4] += e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    tot += t
print(f"one denoiser pass (last of the capture): {len(seq)} launches, {tot / 1e3:.1f} us "
      "(ncu per-launch times: cold cache, serialised, boost clocks - compare SHARES)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} x{a[0]:3d} {a[1] / 1e3:9.1f} us {100 * a[1] / tot:5.1f}%  dram R {a[2] / 1e6:8.1f} MB W {a[3] / 1e6:8.1f} MB "
          f"({(a[2] + a[3]) / a[1]:6.0f} GB/s)  tensor-pipe active {a[4] / a[1]:5.1f}%")
