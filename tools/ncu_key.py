"""Key counters of the first kernel of one or more ncu reports, side by side.
usage: python tools/ncu_key.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum", "sm__inst_executed_pipe_uniform.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__icc_requests.sum", "sm__icc_requests_lookup_miss.sum", "sm__icc_requests_lookup_hit.sum"]
STALL = "smsp__average_warps_issue_stalled_"
cols = {}
names = []
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    cols[path] = d
    names = hdr
allk = KEYS + sorted(k for k in names if k.startswith(STALL) and k.endswith("_per_warp_active.pct") is False and "per_issue" in k and "not_issued" not in k)
icc = sorted(k for k in names if "icc" in k or "instruction_cache" in k.lower())
print("%-72s" % "metric" + "".join("%18s" % p.split("/")[-1].replace(".ncu-rep", "")[-17:] for p in sys.argv[1:]))
for k in allk + [x for x in icc if x not in allk]:
    if any(k in cols[p] for p in sys.argv[1:]):
        print("%-72s" % k[-72:] + "".join("%18s" % cols[p].get(k, "-")[:17] for p in sys.argv[1:]))
