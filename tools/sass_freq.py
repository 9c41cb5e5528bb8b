"""Per-SASS-instruction execution frequency from an ncu report (how many instructions run per warp-substep).
usage: python tools/sass_freq.py report.ncu-rep <warps*substeps per launch>"""
import csv, subprocess, collections, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
WS = float(sys.argv[2])
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []; kern = 0
for r in rows:
    if r and r[0] == "Kernel Name": kern += 1; continue
    if r and r[0] == "Address": hdr = r; iI = hdr.index("Instructions Executed"); continue
    if hdr and kern == 1 and len(r) > iI:
        try: data.append((r[1], int(r[iI])))
        except ValueError: pass
tot = sum(n for _, n in data)
print(len(data), 'SASS instrs; total executed', tot, '; per warp-substep %.1f' % (tot / WS))
for lo, hi in ((0.99, 1e9), (0.5, 0.99), (0.1, 0.5), (0.01, 0.1), (0, 0.01)):
    sel = [(s, n) for s, n in data if lo <= n / WS < hi]
    print('freq [%g,%g): %4d static instrs, %7.1f dynamic instrs per warp-substep' % (lo, hi, len(sel), sum(n for _, n in sel) / WS))
ops = collections.Counter()
for src, n in data:
    t = src.split()
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += n / WS
print('opcode mix (dynamic per warp-substep):', [(k, round(v, 1)) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:24]])
