"""Summarise an .ncu-rep (run where ncu is installed): key metrics, stall reasons, opcode mix, hottest
SASS lines.  usage: ncu_summary.py report.ncu-rep [n_top]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for kr in rows[2:]:
    vals = kr
    get = lambda name: next((vals[i] for i, h in enumerate(hdr) if h == name), None)
    print("kernel:", get("Kernel Name"), " grid", get("launch__grid_size"), " block", get("launch__block_size"))
    for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
              "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]:
        i = next((i for i, h in enumerate(hdr) if h == m), None)
        if i is not None:
            print(f"  {m:75s} {vals[i]:>16s} {units[i]}")
    print("  stall reasons (warps per issue-active cycle):")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try: st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError: pass
    for v, n in sorted(st, reverse=True)[:8]:
        print(f"    {n:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
body = [r for r in rows[2:] if len(r) > 6 and r[5].isdigit()]
tot = sum(int(r[5]) for r in body); samp = sum(int(r[2]) for r in body)
c = Counter()
for r in body:
    t = r[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += int(r[5])
print("warp instructions:", tot, " samples:", samp)
print("opcode mix %:", [(k, round(100 * v / tot, 1)) for k, v in c.most_common(18)])
print("hottest SASS (samples, executed, instruction):")
for r in sorted(body, key=lambda r: -int(r[2]))[:ntop]:
    print(f"  {r[2]:>6s} {r[5]:>10s}  {r[1].strip()[:100]}")
