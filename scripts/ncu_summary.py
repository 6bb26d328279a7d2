"""Summarise an .ncu-rep (read here, no GPU): key metrics + top stall instructions. Usage: ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]
out = []
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        out.append(f"{h} = {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
if hi is not None:
    rows = rows[hi:]
    h = rows[0]
    ci = {n: i for i, n in enumerate(h)}
    key = ci["Warp Stall Sampling (All Samples)"]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    srcc = ci.get("Source")
    inst = ci.get("Instructions Executed")
    tot = sum(float(r[key] or 0) for r in rows[1:] if len(r) > key)
    out.append(f"--- top instructions by stall samples (total {tot:.0f}) ---")
    top = sorted(rows[1:], key=lambda r: -float(r[key] or 0))[:45]
    for r in top:
        why = sorted(((float(r[ci[n]] or 0), n[6:]) for n in stalls), reverse=True)[:2]
        out.append(f"{float(r[key]):8.0f} {100*float(r[key])/tot:5.2f}%  {r[0][-5:]}  {r[srcc].strip()[:70]:70s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}")
    # opcode histogram weighted by executed count
    if inst is not None:
        c = collections.Counter()
        for r in rows[1:]:
            try:
                op = r[srcc].split()[0] if not r[srcc].startswith("@") else r[srcc].split()[1]
                c[op.split(".")[0]] += float(r[inst] or 0)
            except Exception:
                pass
        tt = sum(c.values())
        out.append("--- executed warp-instructions by opcode ---")
        out += [f"{v:14.0f} {100*v/tt:5.1f}%  {k}" for k, v in c.most_common(25)]
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("```\n" + txt + "\n```\n")
