"""Key metrics + hottest source lines of an .ncu-rep (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]
"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(r[idx["Kernel Name"]][:70], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
        for k in KEEP:
            if k in idx:
                print(f"    {k} = {r[idx[k]]} {units[idx[k]]}")
        # warp-state sampling (pc sampling counters of --set full): share of all samples per stall reason
        samp = []
        for h in hdr:
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    samp.append((float(r[idx[h]].replace(",", "")), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in samp) or 1.0
        for v, name in sorted(samp, reverse=True)[:8]:
            print(f"    warp samples {name} = {100 * v / tot:.1f} %")
        print()


def source(path, top):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split("\n\n")
    for blk in blocks:
        rows = list(csv.reader(io.StringIO(blk)))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        try:
            si = hdr.index("Source")
            wi = [i for i, h in enumerate(hdr) if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)"][0]
        except (ValueError, IndexError):
            continue
        data = []
        for r in rows[1:]:
            try:
                data.append((float(r[wi]), r[si]))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1
        print("---- hottest lines (stall samples) ----")
        for s, src in sorted(data, key=lambda t: -t[0])[:top]:
            print(f"{100 * s / tot:5.1f}%  {src.strip()[:140]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if "--source" in sys.argv:
        source(sys.argv[1], int(sys.argv[sys.argv.index("--source") + 1]))
