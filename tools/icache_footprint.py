"""I-cache footprint of the hot code from an ncu report: distinct 128 B lines holding SASS instructions that
execute at least `thr` times per warp-substep.  usage: icache_footprint.py report.ncu-rep <warps*substeps>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
WS = float(sys.argv[2])
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []; kern = 0
for r in rows:
    if r and r[0] == "Kernel Name": kern += 1; continue
    if r and r[0] == "Address": hdr = r; iI = hdr.index("Instructions Executed"); continue
    if hdr and kern == 1 and len(r) > iI:
        try: data.append((int(r[0], 16), int(r[iI])))
        except ValueError: pass
base = min(a for a, _ in data)
print("kernel spans %.1f KB (%d instrs)" % ((max(a for a, _ in data) - base) / 1024.0, len(data)))
for thr in (1.0, 0.5, 0.1, 0.02, 0.001):
    lines = {(a - base) // 128 for a, n in data if n / WS >= thr}
    print("freq >= %-5g per warp-substep: %4d instrs in %4d lines = %.1f KB" %
          (thr, sum(1 for a, n in data if n / WS >= thr), len(lines), len(lines) * 128 / 1024.0))
