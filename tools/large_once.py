"""One fused launch (8 command periods) of 131,072 6v6 worlds after the usual settle: the configs[4] shard, for ncu."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
import torch
W, R, P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 12, 8
b = sim.WorldBatch(W, R, seed=0x5351)
b.gen_commands(0, 0); b.step(120)
for k in range(3):
    b.gen_commands(1 + k, 0); b.step(30)
stage = torch.empty(P * W * R * 32, dtype=torch.uint8, device="cuda")
for k in range(P):
    b.gen_commands_into(stage.data_ptr() + k * W * R * 32, 4 + k, 0)
b.sync(); b.timer_start(); b.step_sequence_device(stage.data_ptr(), P, 30); ms = b.timer_stop()
print("W=%d: %.3f ms, %.4e world-steps/s" % (W, ms, W * P * 30 / (ms * 1e-3)))
