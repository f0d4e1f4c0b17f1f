"""ncu_summary.py <report.ncu-rep> <out.md> [title] -- tracked summary (profiles/) of one `ncu --set full` capture: launch
shape, issue / pipe utilisation, DRAM bytes, stall reasons and the SASS opcode mix of the first kernel in the report.
Runs here (no GPU).  Numbers under a profiler are evidence of WHERE time goes, never bench values."""
import collections, csv, io, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = [f"# {title}", "", f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one B200); summarised by",
         "`scripts/ncu_summary.py`.  Numbers under a profiler are evidence of WHERE time goes, never bench values.", "",
         "| metric | unit | value |", "|---|---|---|"]
for k in keep:
    for i, h in enumerate(hdr):
        if h == k:
            lines.append(f"| {k} | {units[i]} | {data[0][i]} |")
st = [(h, data[0][i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
lines += ["", "## warp stall reasons (average warps stalled per issue-active cycle)", "", "| reason | value |", "|---|---|"]
for h, v in sorted(st, key=lambda t: -float(t[1] or 0)):
    if float(v or 0) > 0.01:
        lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {float(v):.3f} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    shdr = srows[1]
    ix = {h: i for i, h in enumerate(shdr)}
    first = []
    for r in srows[2:]:
        if len(r) == len(shdr) and r[ix["# Samples"]].isdigit():
            first.append(r)
        elif first:
            break
    by, ex = collections.Counter(), collections.Counter()
    for r in first:
        op = re.sub(r"^@!?U?P\d\s+", "", r[ix["Source"]].strip()).split()[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "UTMA", "LDG", "STG", "LDS", "STS", "ATOMS", "RED", "SYNCS", "UBLKCP")) else op.split(".")[0]
        by[op] += int(r[ix["# Samples"]])
        ex[op] += int(r[ix["Instructions Executed"]])
    tot, tex = max(1, sum(by.values())), max(1, sum(ex.values()))
    lines += ["", "## SASS opcode mix of the launch (source page)", "",
              "| opcode | % of stall samples | % of executed warp instructions |", "|---|---|---|"]
    for op, s in ex.most_common(22):
        lines.append(f"| {op} | {100 * by[op] / tot:.1f} | {100 * ex[op] / tex:.1f} |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
