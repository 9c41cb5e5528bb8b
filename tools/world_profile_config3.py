"""Per-world counters (-DSSLSIM_PROFILE build) for configs[2]: 11v11 worlds on the Division-A field, crowded
penalty-area starts, kicks and dribbling; one fused launch of 10 command periods."""
import ctypes as C, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
import torch
L = sim.lib()
L.world_batch_profile_counters.argtypes = [C.c_void_p, C.c_void_p]
W, R, P = 16384, 22, 10
WARM = float(os.environ.get("SSLSIM_WARM", "0.85"))   # warmstart_factor; 0 = cold-start kernels
b = sim.WorldBatch(W, R, fld="FIELD_DIV_A", init_mode=1, params=sim.default_params(warmstart_factor=WARM))
print("warmstart_factor", WARM)
for p in range(4):
    b.gen_commands(p, 1); b.step(30)
stage = torch.empty(P * W * R * 32, dtype=torch.uint8, device="cuda")
for k in range(P):
    b.gen_commands_into(stage.data_ptr() + k * W * R * 32, 4 + k, 1)
out = np.zeros((W, 32), np.uint32)
L.world_batch_profile_counters(b._h, None)
b.sync(); b.timer_start(); b.step_sequence_device(stage.data_ptr(), P, 30); ms = b.timer_stop()
L.world_batch_profile_counters(b._h, out.ctypes.data_as(C.c_void_p))
cyc = out[:, 0].astype(np.float64)
print("fused launch %.3f ms (%d substeps, %.3e world-steps/s); cycles mean %.0f p50 %.0f p90 %.0f max %.0f" %
      (ms, P * 60, W * P * 30 / (ms * 1e-3), cyc.mean(), np.percentile(cyc, 50), np.percentile(cyc, 90), cyc.max()))
names = ["cycles", "generic_substeps", "pair_substeps", "iterations", "levels", "pair_rows", "static_substeps"]
print("means:", " ".join("%s=%.1f" % (n, v) for n, v in zip(names, out[:, :7].mean(0))))
sec = ["actuation", "row_setup", "broadphase", "pair_rows", "solve", "rest"]
print("mean per substep: " + " ".join("%s=%.0f" % (n, v / (P * 60)) for n, v in zip(sec, out[:, 8:14].mean(0))))
cfgn = ["ground only", "pair", "walls", "pair+walls", "boxes", "pair+boxes", "walls+boxes", "pair+walls+boxes"]
tc = out[:, 16:24].astype(np.float64).sum(0); ti = out[:, 24:32].astype(np.float64).sum(0)
for c in range(8):
    if ti[c]: print("  %-18s cycles %5.1f%%  iterations %5.1f%%  cycles/iteration %6.0f" % (cfgn[c], 100 * tc[c] / tc.sum(), 100 * ti[c] / ti.sum(), tc[c] / ti[c]))
why = np.bitwise_or.reduce(out[:, 7])
print("generic-path reasons seen (1 deep ground,2 ceil,8 deep wall,16 >2 boxes,32 deep box,64 deep pair): %d" % why)
