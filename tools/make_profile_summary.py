"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep, bench_*.json} into the tracked summaries under profiles/.
usage: python tools/make_profile_summary.py <tag> [ncu-rep] [launches.csv] [bench.json] [world-steps per launch] [worlds] [command]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out/prof_r1.ncu-rep")
launches = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out/launches.csv")
bench = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out/bench_r1.json")
WSPL = int(sys.argv[5]) if len(sys.argv) > 5 else 4096 * 30 * 10   # worlds x world-steps of one captured launch
WORLDS = int(sys.argv[6]) if len(sys.argv) > 6 else 4096
CMD = sys.argv[7] if len(sys.argv) > 7 else "python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --large-worlds 0"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# 1. launch list: per-kernel totals and shares
rows = list(csv.reader(open(launches)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
kn, mv = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
agg = {}
for r in rows[h + 1:]:
    if len(r) > mv:
        try:
            agg.setdefault(r[kn], []).append(float(r[mv].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, tag + "_launches.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none: %s (%d 6v6 worlds; the fused pass runs %d "
            "world-steps per step_worlds launch)\n" % (CMD, WORLDS, WSPL // WORLDS))
    f.write("per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes\n\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-70s launches %3d  total %9.1f us  avg %8.1f us  share %5.1f%%\n" %
                (k[:70], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
subprocess.run(["cp", launches, os.path.join(out, tag + "_launches.csv")])

# 2. step_worlds --set full summary
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(os.path.join(out, tag + "_step_worlds_ncu.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:step_worlds: %s\n" % CMD)
    for i, hname in enumerate(hdr):
        if hname in keep or ("issue_stalled" in hname and hname.endswith("per_issue_active.ratio")):
            f.write("%-82s %-16s %s\n" % (hname, units[i], " | ".join(r[i] for r in data)))
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
def tobytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
tr = sum(tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]) for r in data) / len(data)
ii = hdr.index("smsp__inst_executed.sum")
inst = sum(float(r[ii].replace(",", "")) for r in data) / len(data)
json.dump({"dram_bytes_per_launch": tr, "warp_instructions_per_launch": inst,
           "world_steps_per_launch": WSPL, "launch_shape": "%d 6v6 worlds x %d world-steps" % (WORLDS, WSPL // WORLDS),
           "kernel": "step_worlds", "source": os.path.basename(rep), "capture": tag,
           "how": "dram__bytes_read.sum + dram__bytes_write.sum, mean over %d captured launches, ncu --set full" % len(data)},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)

# 3. per-line / footprint breakdowns
for tool, name in (("ncu_lines.py", "_step_worlds_lines.txt"), ("icache_footprint.py", "_step_worlds_icache.txt"),
                   ("sass_freq.py", "_step_worlds_sass_freq.txt")):
    # warp-substeps per launch = (worlds / 2 warps) x (2 substeps per step) x steps = worlds x steps
    args = [sys.executable, os.path.join(ROOT, "tools", tool), rep] + (["40"] if tool == "ncu_lines.py" else [str(WSPL)])
    txt = subprocess.run(args, capture_output=True, text=True).stdout
    open(os.path.join(out, tag + name), "w").write(txt)
if os.path.exists(bench):
    subprocess.run(["cp", bench, os.path.join(out, tag + "_bench.json")])
print("wrote profiles/%s_*" % tag, "traffic per launch %.2f MB" % (tr / 1e6))
