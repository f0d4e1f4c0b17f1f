"""Turns gpurun_out/prof_sweep.ncu-rep + launches.csv into the tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = os.path.join(ROOT, "gpurun_out", "prof_sweep.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ["Kernel Name", "gpu__time_duration.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = [f"# ncu --set full summary of the dominant kernel ({tag})", "",
         "Source: `gpurun_out/prof_sweep.ncu-rep` captured by `scripts/gpu_profile.sh` (ncu --set full --clock-control none",
         "--import-source on -k regex:sweep_philox) on one B200; numbers under a profiler are evidence of WHERE time goes,",
         "never bench values.", "", "| metric | unit | value |", "|---|---|---|"]
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
shdr = srows[1]
ix = {h: i for i, h in enumerate(shdr)}
blk = [r for r in srows[2:] if len(r) == len(shdr) and r[ix["# Samples"]].isdigit()]
first = []
for r in srows[2:]:
    if len(r) == len(shdr) and r[ix["# Samples"]].isdigit():
        first.append(r)
    elif first:
        break
by, ex = collections.Counter(), collections.Counter()
for r in first:
    op = re.sub(r"^@!?U?P\d\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
    by[op] += int(r[ix["# Samples"]])
    ex[op] += int(r[ix["Instructions Executed"]])
tot, tex = sum(by.values()), sum(ex.values())
lines += ["", "## SASS opcode mix of the launch (source page, --import-source on)", "",
          "| opcode | % of stall samples | % of executed warp instructions |", "|---|---|---|"]
for op, s in by.most_common(16):
    lines.append(f"| {op} | {100 * s / tot:.1f} | {100 * ex[op] / tex:.1f} |")
open(os.path.join(out_dir, f"{tag}_sweep_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
import json
def _val(name):
    i = hdr.index(name)
    v = float(data[0][i].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
series = int(sys.argv[2]) if len(sys.argv) > 2 else 11   # store intervals per launch of the profiled bench.py run
inst = float(data[0][hdr.index("smsp__inst_executed.sum")].replace(",", ""))
def _f(name):
    return float(data[0][hdr.index(name)].replace(",", ""))
# FP64 work the kernel really executes: thread-level DFMA (2 flop) + DMUL + DADD per elapsed cycle x elapsed cycles
cycles = _f("sm__cycles_elapsed.max")
flop = (2 * _f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") +
        _f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") +
        _f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")) * cycles
json.dump({"chains": 1 << 27, "mc_steps": 10, "series": series,
           "warp_inst_per_warp_step": inst / ((1 << 27) / 32 * 10 * series),
           "issue_active_pct": float(data[0][hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
           "fp64_flop_per_chain_step": flop / ((1 << 27) * 10 * series),
           "fp64_pipe_pct": _f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "dram_bytes_per_launch": _val("dram__bytes_read.sum") + _val("dram__bytes_write.sum"),
           "source": f"profiles/{tag}_sweep_ncu_summary.md (ncu --set full, bench.py default shape)"},
          open(os.path.join(out_dir, "traffic.json"), "w"))

# launch list: per-kernel totals and shares
lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lc):
    txt = [l for l in open(lc) if not l.startswith("==")]
    rr = list(csv.DictReader(io.StringIO("".join(txt))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rr:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            u = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
            name = re.sub(r"\(.*", "", r["Kernel Name"])
            agg[name][0] += 1
            agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    out = [f"# launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 32 --warmup 16`", "",
           "Per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes.", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
        out.append(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
    open(os.path.join(out_dir, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    open(os.path.join(out_dir, f"{tag}_launches.csv"), "w").write("".join(txt))
print("wrote", out_dir)
