"""Throughput of the other BASELINE.json configurations (parity-test cases, not the bench line), for BASELINE.md."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sim = importlib.import_module("ssl-sim_b200")

def fused(W, R, fld, mode, kind, periods=10, settle=4):
    b = sim.WorldBatch(W, R, fld=fld, init_mode=mode)
    for p in range(settle):
        b.gen_commands(p, kind); b.step(30)
    stage = torch.empty(periods * W * R * 32, dtype=torch.uint8, device="cuda")
    for k in range(periods):
        b.gen_commands_into(stage.data_ptr() + k * W * R * 32, settle + k, kind)
    b.sync(); b.timer_start(); b.step_sequence_device(stage.data_ptr(), periods, 30); ms = b.timer_stop()
    b.close()
    return W * periods * 30 / (ms * 1e-3)

# configs[2]: 65,536 x 11v11, crowded penalty-area starts, Division-A field
print("config 3 (65,536 x 11v11 crowded, kicks+dribble): %.3e world-steps/s" % fused(65536, 22, "FIELD_DIV_A", 1, 1))
# configs[3]: 16,384 x 6v6, kicker/dribbler every step, detection gather for 64 observed worlds every step
W, R = 16384, 12
b = sim.WorldBatch(W, R)
for p in range(4):
    b.gen_commands(p, 1); b.step(30)
ids = np.arange(64, dtype=np.int32) * 256
steps = 600
b.sync(); t0 = time.perf_counter()
for s in range(steps):
    if s % 30 == 0:
        b.gen_commands(4 + s // 30, 1)
    b.step(1)
    det = b.gather_detections(ids)        # K3 + D2H into host memory, synchronises
dt = time.perf_counter() - t0
print("config 4 (16,384 x 6v6, kick/dribble, 64-world detection gather EVERY step, wall clock): %.3e world-steps/s, "
      "%.1f us per step+gather" % (W * steps / dt, dt / steps * 1e6))
print("config 4 without the per-step gather, fused: %.3e world-steps/s" % fused(16384, 12, "FIELD_2015", 0, 1))
print("config 5 shard (131,072 x 6v6), fused: %.3e world-steps/s" % fused(131072, 12, "FIELD_2015", 0, 0))
