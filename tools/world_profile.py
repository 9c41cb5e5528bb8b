"""Per-world cost counters from a -DSSLSIM_PROFILE build (make -C ssl-sim_b200/csrc clean all EXTRA=-DSSLSIM_PROFILE).
Shows what the slowest worlds of a launch are doing (the kernel ends when its slowest warp does)."""
import ctypes as C, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
L = sim.lib()
L.world_batch_profile_counters.argtypes = [C.c_void_p, C.c_void_p]
W, R = 4096, 12
b = sim.WorldBatch(W, R)
b.gen_commands(0, 0); b.step(120)
for p in range(1, int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    b.gen_commands(p, 0); b.step(30)
out = np.zeros((W, 32), np.uint32)
L.world_batch_profile_counters(b._h, None)     # allocates the counter buffer
b.gen_commands(1000, 0); b.timer_start(); b.step(30); ms = b.timer_stop()
L.world_batch_profile_counters(b._h, out.ctypes.data_as(C.c_void_p))
cyc = out[:, 0].astype(np.float64)
print("launch %.3f ms; cycles per world-group: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %.0f" %
      (ms, cyc.mean(), np.percentile(cyc, 50), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max()))
names = ["cycles", "generic_substeps", "pair_substeps", "iterations", "levels", "pair_rows", "static_substeps", "why(1 deep ground,2 ceil,4 both walls,8 deep wall,16 two boxes,32 deep box,64 deep pair)"]
order = np.argsort(-cyc)
print("slowest 12 worlds (of 60 substeps):")
for w in order[:12]:
    print("  world %5d " % w + " ".join("%s=%d" % (n, v) for n, v in zip(names, out[w])))
print("means:", " ".join("%s=%.1f" % (n, v) for n, v in zip(names, out.mean(0))))
for k, n in enumerate(names[1:7], 1):
    print("corr(cycles, %s) = %.2f" % (n, np.corrcoef(cyc, out[:, k])[0, 1]))
