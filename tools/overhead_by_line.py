"""Where the control-flow and move instructions are: joins the per-address execution counts of an ncu report with
the source lines nvdisasm attributes to each SASS instruction of step_worlds<16,false,true>, and lists the source
lines with the most executed BRA/BSSY/BSYNC and MOV/IMAD.MOV/HFMA2 instructions per warp-substep.
usage: python tools/overhead_by_line.py report.ncu-rep [warp-substeps per launch]"""
import csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]; WS = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 300.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "ssl-sim_b200/csrc/step_kernel.o")], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and "Li16ELb0ELb1" in l)
line_of = {}; cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//---"):
        if line_of: break
        continue
    m = re.search(r'//## File ".*step_kernel.cu", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
s = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
e = next((i for i in range(s + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
h = rows[s]; A, S, E = h.index("Address"), h.index("Source"), h.index("Instructions Executed")
base = int(rows[s + 1][A], 16)
ctl, mov, allc = {}, {}, {}
for r in rows[s + 1:e]:
    t = r[S].strip(); op = (t.split()[1] if t.startswith("@") else t.split()[0])
    ln = line_of.get(int(r[A], 16) - base); ex = float(r[E]) / WS
    allc[ln] = allc.get(ln, 0) + ex
    if op.split(".")[0] in ("BRA", "BSSY", "BSYNC", "BREAK"): ctl[ln] = ctl.get(ln, 0) + ex
    if op.split(".")[0] in ("MOV", "HFMA2", "CS2R", "UMOV") or op.startswith("IMAD.MOV"): mov[ln] = mov.get(ln, 0) + ex
src = open(os.path.join(ROOT, "ssl-sim_b200/csrc/step_kernel.cu")).read().splitlines()
for name, d in (("control flow (BRA/BSSY/BSYNC)", ctl), ("moves (MOV/IMAD.MOV/HFMA2)", mov)):
    print("%s: %.1f per warp-substep; top source lines:" % (name, sum(d.values())))
    for ln, v in sorted(d.items(), key=lambda kv: -kv[1])[:22]:
        print("  %5.2f  (all instrs of the line %6.2f)  L%s %s" % (v, allc.get(ln, 0), ln, src[ln - 1].strip()[:90] if ln else ""))
