"""Hottest SASS instructions (warp-stall samples) of each kernel in an .ncu-rep.

    python tools/ncu_source.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, data=[])
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["data"].append(r)
STALLS = ["stall_barrier", "stall_lg", "stall_long_sb", "stall_math", "stall_mio", "stall_short_sb", "stall_wait",
          "stall_not_selected", "stall_sleep", "stall_tex", "stall_membar", "stall_dispatch", "stall_branch_resolving"]
for k in kernels:
    h = {n: i for i, n in enumerate(k["hdr"])}
    si, ai = h["Source"], h["Warp Stall Sampling (All Samples)"]
    tot = sum(float(r[ai] or 0) for r in k["data"]) or 1
    print("====", k["name"][:100], f"({len(k['data'])} SASS instructions, {int(tot)} samples)")
    agg = {s: sum(float(r[h[s]] or 0) for r in k["data"]) for s in STALLS if s in h}
    print("   stall mix:", ", ".join(f"{s[6:]} {100 * v / tot:.0f}%" for s, v in sorted(agg.items(), key=lambda t: -t[1])[:7]))
    for r in sorted(k["data"], key=lambda r: -float(r[ai] or 0))[:top]:
        why = max(((s, float(r[h[s]] or 0)) for s in STALLS if s in h), key=lambda t: t[1])
        print(f"  {100 * float(r[ai] or 0) / tot:5.1f}%  {r[si].strip()[:90]:90s} {why[0][6:]}")
