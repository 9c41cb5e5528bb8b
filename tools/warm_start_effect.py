"""What warm starting (docs/STEP_SPEC.md 4.5a, spec v3) changes against the cold-start step of spec v2: Bullet warm-starts
every contact point that persisted from the previous substep at 0.85 x its last impulses;
spec v2 started every row from zero (its deviation D2, closed in v3; warmstart_factor = 0 still gives that step).

This script follows the WARM-started trajectory (the default parameters) on the CPU oracle and, at every world-step, also
takes the same step COLD (warmstart_factor = 0) from the same state, and reports how far the two results are apart after
that one step -- the per-step deviation north_star's tolerance (1e-4 m, 1e-3 rad per step) is stated in.  (Comparing two
free-running trajectories would only measure chaos: robot contacts amplify a last-bit difference to centimetres within
seconds.)  profiles/r2_warm_start_effect.txt is its output.

usage: python tools/warm_start_effect.py [config2|config3] [worlds] [steps]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob  # noqa: E402

CFG = sys.argv[1] if len(sys.argv) > 1 else "config2"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 64
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 900
L = ob.lib()

if CFG == "config3":
    R, kind = 22, 1
    w = ob.Worlds(W, R, fld=ob.FIELD_DIV_A, mode=1)
else:
    R, kind = 12, 0
    w = ob.Worlds(W, R)
# the cold step = the same oracle with warmstart_factor 0 (the step of spec v2); `w` runs the default (spec v3, 0.85)
cold = ob.Worlds(1, R, fld=(ob.FIELD_DIV_A if CFG == "config3" else ob.FIELD_2015), params=ob.default_params(warmstart_factor=0.0))
one = ob.Worlds(1, R, fld=(ob.FIELD_DIV_A if CFG == "config3" else ob.FIELD_2015))
dpos, dvel, dyaw, rows_per_substep, differing, total = [], [], [], [], 0, 0
ev_diff = 0
for step in range(STEPS):
    if step % 30 == 0:
        w.gen_commands(step // 30, kind)
    for i in range(W):
        # cold step (the spec) from the same state
        cold.pos[0] = w.pos[i]; cold.vel[0] = w.vel[i]; cold.aux[0] = w.aux[i]
        cold.cmd = w.cmd[i:i + 1].copy()
        cold.step(1)
        c0 = int(w.aux[i, 7])
        one.pos[0] = w.pos[i]; one.vel[0] = w.vel[i]; one.aux[0] = w.aux[i]; one.warm[0] = w.warm[i]
        one.cmd = w.cmd[i:i + 1].copy()
        one.step(1)
        w.pos[i] = one.pos[0]; w.vel[i] = one.vel[0]; w.aux[i] = one.aux[0]; w.warm[i] = one.warm[0]
        if step < 120 and CFG != "config3":
            continue          # the drop from 0.5 m
        rows_per_substep.append((int(w.aux[i, 7]) - c0) / 2.0)
        d = np.abs(w.pos[i, :, :3].astype(np.float64) - cold.pos[0, :, :3]).max()
        dv = np.abs(w.vel[i, :, :3].astype(np.float64) - cold.vel[0, :, :3]).max()
        dy = np.abs(w.pos[i, :R, 3].astype(np.float64) - cold.pos[0, :R, 3]).max()
        dpos.append(d); dvel.append(dv); dyaw.append(dy)
        total += 1
        differing += d > 0 or dv > 0
        ev_diff += int(((int(w.aux[i, 0]) ^ int(cold.aux[0, 0])) & 0x7000) != 0) + int(w.aux[i, 2] != cold.aux[0, 2])
dpos, dvel, dyaw = np.array(dpos), np.array(dvel), np.array(dyaw)
print("%s workload on the oracle: %d worlds x %d world-steps, %d one-step comparisons (warm-started step vs the spec's cold step from the same state)"
      % (CFG, W, STEPS, total))
print("contact rows per substep and world: mean %.1f" % np.mean(rows_per_substep))
print("steps whose result differs at all: %.1f %%" % (100.0 * differing / total))
for name, a, tol, unit in (("position", dpos, 1e-4, "m"), ("velocity", dvel, None, "m/s"), ("yaw", dyaw, 1e-3, "rad")):
    q = np.quantile(a, [0.5, 0.99, 0.999])
    line = "%-9s per-step difference: median %.2e, 99 %% %.2e, 99.9 %% %.2e, max %.2e %s" % (name, q[0], q[1], q[2], a.max(), unit)
    if tol:
        line += "   (north_star tolerance %.0e: exceeded in %.3f %% of the steps)" % (tol, 100.0 * (a > tol).mean())
    print(line)
print("steps whose goal / out events or touch mask differ: %d" % ev_diff)
