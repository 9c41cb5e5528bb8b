"""Scratch timing of one build of the library: fused 10-period launches at 4,096 and 131,072 6v6 worlds (the bench's
device arm without its bookkeeping).  Used by tools/variant_run.sh for A/B/C comparisons inside one gpurun session."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
import torch

WARM = float(os.environ.get("SSLSIM_WARM", "0.85"))   # warmstart_factor; 0 = the cold-start kernels of spec v2

CFG = os.environ.get("SSLSIM_CFG", "")   # "3": 11v11 crowded Division-A worlds with kicks and dribbling (size / 2 worlds)

def run(W, R=12, P=10, reps=4, settle_periods=40):
    kind = 0
    if CFG == "3":
        R, kind, W = 22, 1, W // 2
        b = sim.WorldBatch(W, R, fld="FIELD_DIV_A", init_mode=1, params=sim.default_params(warmstart_factor=WARM))
        settle_periods = 4
    else:
        kind = 1 if CFG == "4" else 0
        b = sim.WorldBatch(W, R, params=sim.default_params(warmstart_factor=WARM))
    b.gen_commands(0, kind); b.step(120)
    for p in range(1, settle_periods):
        b.gen_commands(p, kind); b.step(30)
    stage = torch.empty(P * W * R * 32, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for r in range(reps + 1):
        for k in range(P):
            b.gen_commands_into(stage.data_ptr() + k * W * R * 32, 1000 + r * P + k, kind)
        b.sync(); b.timer_start(); b.step_sequence_device(stage.data_ptr(), P, 30); ms = b.timer_stop()
        if r: tot += ms
    st = b.reduce_stats()
    b.close()
    return W * P * 30 * reps / (tot * 1e-3), st["checksum"]

if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 131072]
    out = []
    for W in sizes:
        v, ck = run(W, reps=8 if W <= 16384 else 3)
        out.append("W=%d %.4e" % (W, v))
    print(tag, "cfg=%s warm=%g" % (CFG or "2/5", WARM), " | ".join(out), "ck", ck, flush=True)
