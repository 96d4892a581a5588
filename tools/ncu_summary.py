"""Summarise an .ncu-rep (run where ncu is installed): key metrics + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print("kernel:", name[:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {r[i]:>16s} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warp-cycles per issued instr):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, f = [], None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        f = r[1].split("/")[-1]
    elif len(r) > 8 and r[2] == "-" and r[0].isdigit():
        try:
            agg.append((int(r[6]), int(r[7]), f, int(r[0]), r[1].strip()[:95]))
        except ValueError:
            pass
ts, te = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print(f"hottest source lines by stall samples (total samples {ts}, instructions {te}):")
for s_, e_, f, l, t in sorted(agg, reverse=True)[:top]:
    print(f"  {100 * s_ / ts:5.1f}% samp {100 * e_ / te:5.1f}% inst  {f}:{l}  {t}")
