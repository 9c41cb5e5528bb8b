"""Where the instruction-fetch stalls are: per SASS instruction stall_no_inst / stall_branch_resolving samples from an
ncu --set full --import-source on report.  usage: python tools/ncu_noinst.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first kernel only
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(lines[start:end]))
h = rows[0]; rows = rows[1:]
ci = {n: h.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "stall_no_inst", "stall_branch_resolving", "stall_wait", "stall_short_sb", "stall_selected")}
def f(r, n):
    try: return float(r[ci[n]])
    except ValueError: return 0.0
tot = {n: sum(f(r, n) for r in rows) for n in ci if n not in ("Address", "Source")}
print("totals:", {k: int(v) for k, v in tot.items()})
base = int(rows[0][ci["Address"]], 16)
print("top instructions by stall_no_inst samples (offset, samples, executed, prev instr):")
order = sorted(range(len(rows)), key=lambda i: -f(rows[i], "stall_no_inst"))[:top]
for i in order:
    r = rows[i]
    print("  +%05x no_inst %6d  exec %9d  %-44s | prev: %s" % (int(r[ci["Address"]], 16) - base, f(r, "stall_no_inst"), f(r, "Instructions Executed"),
          r[ci["Source"]].strip()[:44], rows[i - 1][ci["Source"]].strip()[:40] if i else ""))
# no_inst samples by 128-byte line position: first instruction of a line vs others
first = sum(f(r, "stall_no_inst") for r in rows if (int(r[ci["Address"]], 16) % 128) == 0)
print("no_inst samples on the first instruction of a 128 B line: %.1f%%" % (100 * first / max(tot["stall_no_inst"], 1)))
after_branch = sum(f(rows[i], "stall_no_inst") for i in range(1, len(rows)) if any(k in rows[i - 1][ci["Source"]] for k in ("BRA", "CALL", "RET", "BSYNC", "EXIT")))
print("no_inst samples on an instruction that follows a BRA/CALL/RET/BSYNC in address order: %.1f%%" % (100 * after_branch / max(tot["stall_no_inst"], 1)))
