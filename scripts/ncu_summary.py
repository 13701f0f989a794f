"""Summarise an .ncu-rep (raw + source pages) into text: python scripts/ncu_summary.py rep [units]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None  # e.g. chain-steps in the profiled launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit_row, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_write.sum", "dram__bytes_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for i, h in enumerate(hdr):
    if h in want:
        print(f"  {h} = {vals[i]} {unit_row[i]}")
if units:
    print(f"  warp-instructions per unit = {float(vals[hdr.index('smsp__inst_executed.sum')].replace(',', '')) / units:.1f}")
stall = [(h, float(vals[i].replace(',', '') or 0)) for i, h in enumerate(hdr)
         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for h, v in sorted(stall, key=lambda t: -t[1])[:8]:
    print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
ia, isrc, ist = h2.index("Instructions Executed"), h2.index("Source"), h2.index("Warp Stall Sampling (All Samples)")
tot = collections.Counter()
samples = collections.Counter()
n = 0
for r in rows[2:]:
    try:
        c = int(r[ia])
    except Exception:
        continue
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    tot[op] += c
    samples[op] += int(r[ist] or 0)
    n += c
ns = sum(samples.values()) or 1
print("  opcode mix (executed warp-instructions, stall samples):")
for op, c in tot.most_common(16):
    extra = f"  per-unit {c / units:.1f}" if units else ""
    print(f"    {op:8s} {c / n * 100:5.1f}%  samples {samples[op] / ns * 100:5.1f}%{extra}")
