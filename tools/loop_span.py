"""Byte span of the solver iteration loop of step_worlds<16,false,true> in the built object: from the first pair-row SHFL
of the loop to the loop-closing VOTE.ANY, plus where level_sweep / goal-row code sits (LDS.U8 = level table read)."""
import subprocess, re, sys
o = subprocess.run(["cuobjdump", "-sass", sys.argv[1] if len(sys.argv) > 1 else "ssl-sim_b200/csrc/step_kernel.o"], capture_output=True, text=True).stdout
blk = o.split("Function : ")
k = next(b for b in blk if "Li16ELb0ELb1" in b.split("\n")[0])
ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", k)
print("kernel instructions:", len(ins), "bytes:", len(ins) * 16)
for a, t in ins:
    if "LDS.U8" in t or "VOTE.ANY" in t or "BRA" in t and int(a, 16) > 0x8000 and int(a,16) < 0xb000:
        pass
votes = [int(a, 16) for a, t in ins if t.strip().startswith("VOTE.ANY") or "VOTE.ANY" in t]
lds8 = [int(a, 16) for a, t in ins if "LDS.U8" in t]
print("VOTE.ANY at", [hex(v) for v in votes])
print("LDS.U8 (level table) at", [hex(v) for v in lds8])
# backward branches = loops
for a, t in ins:
    m = re.search(r"BRA\s+(?:U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < int(a, 16):
        print("loop: %s -> %s  span %d B  (%s)" % (a, m.group(1), int(a, 16) - int(m.group(1), 16), t.strip()[:40]))
