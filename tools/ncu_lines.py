"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [kernel-id] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}
hdr = None
kern = 0
for r in rows:
    if r and r[0] == "Function Name":
        kern += 1
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= iI or kern > 1:
        continue
    if r[2] == "-" and r[0].isdigit():   # a source-line summary row
        try:
            agg[int(r[0])] = (int(r[iI]), int(r[iS]), r[1])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total inst %d samples %d" % (tot, tots))
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  L%-4d %s" % (100.0 * v[0] / tot, 100.0 * v[1] / max(tots, 1), ln, v[2][:120]))

if len(sys.argv) > 3:
    # region summary: "name:lo-hi,name:lo-hi"
    print("--- regions")
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(v[0] for ln, v in agg.items() if lo <= ln <= hi); s = sum(v[1] for ln, v in agg.items() if lo <= ln <= hi)
        print("%-22s L%d-%d: %5.2f%% inst %5.2f%% samples" % (name, lo, hi, 100.0 * i / tot, 100.0 * s / max(tots, 1)))
