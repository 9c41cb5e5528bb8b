"""Per-world counters over one fused launch of 10 command periods (-DSSLSIM_PROFILE build)."""
import ctypes as C, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
import torch
L = sim.lib()
L.world_batch_profile_counters.argtypes = [C.c_void_p, C.c_void_p]
W, R, P = 4096, 12, 10
WARM = float(os.environ.get("SSLSIM_WARM", "0.85"))   # warmstart_factor; 0 = cold-start kernels
b = sim.WorldBatch(W, R, params=sim.default_params(warmstart_factor=WARM))
print("warmstart_factor", WARM)
b.gen_commands(0, 0); b.step(120)
for p in range(1, 120):
    b.gen_commands(p, 0); b.step(30)
stage = torch.empty(P * W * R * 32, dtype=torch.uint8, device="cuda")
for k in range(P):
    b.gen_commands_into(stage.data_ptr() + k * W * R * 32, 1000 + k, 0)
out = np.zeros((W, 32), np.uint32)
L.world_batch_profile_counters(b._h, None)
b.sync(); b.timer_start(); b.step_sequence_device(stage.data_ptr(), P, 30); ms = b.timer_stop()
L.world_batch_profile_counters(b._h, out.ctypes.data_as(C.c_void_p))
names = ["cycles", "generic_substeps", "pair_substeps", "iterations", "levels", "pair_rows", "static_substeps", "why"]
cyc = out[:, 0].astype(np.float64)
print("fused launch %.3f ms (%d substeps); cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" %
      (ms, P * 60, cyc.mean(), np.percentile(cyc, 50), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max()))
sec = ["actuation", "row_setup", "broadphase", "pair_rows", "solve", "rest"]
for w in np.argsort(-cyc)[:10]:
    print("  world %5d " % w + " ".join("%s=%d" % (n, v) for n, v in zip(names[:7], out[w])))
    print("              per substep: " + " ".join("%s=%d" % (n, v / (P * 60)) for n, v in zip(sec, out[w, 8:14])))
print("mean per substep: " + " ".join("%s=%.0f" % (n, v / (P * 60)) for n, v in zip(sec, out[:, 8:14].mean(0))))
print("means:", " ".join("%s=%.1f" % (n, v) for n, v in zip(names[:7], out[:, :7].mean(0))))
cfgn = ["ground only", "pair", "walls", "pair+walls", "boxes", "pair+boxes", "walls+boxes", "pair+walls+boxes"]
ev = out[::2]   # one row per warp (two worlds of a warp carry the same counters)
print("solve loop by warp configuration (all warps): share of solve cycles, iterations, cycles per iteration")
tc = ev[:, 16:24].astype(np.float64).sum(0); ti = ev[:, 24:32].astype(np.float64).sum(0)
for c in range(8):
    if ti[c]: print("  %-18s cycles %5.1f%%  iterations %5.1f%%  cycles/iteration %6.0f" % (cfgn[c], 100 * tc[c] / tc.sum(), 100 * ti[c] / ti.sum(), tc[c] / ti[c]))
w = int(np.argmax(cyc))
print("slowest world %d:" % w)
for c in range(8):
    if out[w, 24 + c]: print("  %-18s iterations %5d  cycles/iteration %6.0f" % (cfgn[c], out[w, 24 + c], out[w, 16 + c] / out[w, 24 + c]))
