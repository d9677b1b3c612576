"""Turn ncu exports into the committed summaries under profiles/.

    python tools/make_profile_summary.py <round-tag> <launches.csv> <raw.csv> [microbench.jsonl] [bench.json]

launches.csv : ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv log
raw.csv      : ncu -i prof.ncu-rep --page raw --csv
Writes profiles/<tag>_launches.csv (copy), profiles/<tag>_kernels.md (share of the step per kernel + key counters).
"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, raw = sys.argv[1], sys.argv[2], sys.argv[3]
micro = sys.argv[4] if len(sys.argv) > 4 else None
bench = sys.argv[5] if len(sys.argv) > 5 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))

SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "ms": 1e6, "us": 1e3, "ns": 1, "second": 1e9}
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
lines = [ln for ln in open(launches) if not ln.startswith("==")]
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("rsa::", "")
    v = float(row["Metric Value"].replace(",", "")) * SCALE.get(row["Metric Unit"], 1)
    agg[k][row["Metric Name"]] += v
    if row["Metric Name"] == "gpu__time_duration.sum":
        cnt[k] += 1
ours = {k: a for k, a in agg.items() if not k.startswith("bench")}
tot = sum(a["gpu__time_duration.sum"] for a in ours.values())
md = [f"# {tag}: kernels of one `bench.py --pairs 262144 --steps 2 --warmup 3` run under ncu", "",
      "Per-launch times from `ncu --metrics gpu__time_duration.sum --clock-control none` are cold-cache and",
      "serialised: compare SHARES, not absolutes (see bench.py for the CUDA-event numbers).", "",
      "| kernel | launches | mean ms | share of our kernels | DRAM read MB/launch | DRAM write MB/launch |", "|---|---|---|---|---|---|"]
for k, a in sorted(ours.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    n = cnt[k]
    md.append(f"| `{k}` | {n} | {a['gpu__time_duration.sum']/n/1e6:.3f} | {a['gpu__time_duration.sum']/tot:.3f} | "
              f"{a['dram__bytes_read.sum']/n/1e6:.1f} | {a['dram__bytes_write.sum']/n/1e6:.1f} |")
md += ["", "## `ncu --set full` counters (one launch each)", ""]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
for vals in rows[2:]:
    md += [f"### `{vals[ki].split('(')[0]}`", "", "| counter | value | unit |", "|---|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            md.append(f"| {w} | {vals[i]} | {units[i]} |")
    md.append("")
if micro and os.path.exists(micro):
    shutil.copy(micro, os.path.join(out_dir, f"{tag}_dpx_microbench.jsonl"))
    md += ["## Issue-rate microbenchmark (tools/dpx_microbench, same box)", "", "```"] + [ln.rstrip() for ln in open(micro)] + ["```", ""]
if bench and os.path.exists(bench):
    shutil.copy(bench, os.path.join(out_dir, f"{tag}_bench.json"))
    md += ["## bench.py line of the same build (not under ncu)", "", "```json", json.dumps(json.load(open(bench)), indent=1), "```", ""]
open(os.path.join(out_dir, f"{tag}_kernels.md"), "w").write("\n".join(md))
print("\n".join(md[:14]))
