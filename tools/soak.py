"""Soak / scale check: long fused runs must stay finite and overflow-free, and a 1,048,576-world batch must fit and run."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sim = importlib.import_module("ssl-sim_b200")
def run(W, R, periods, kind, fld="FIELD_2015", mode=0):
    b = sim.WorldBatch(W, R, fld=fld, init_mode=mode)
    P = 50
    stage = torch.empty(P * W * R * 32, dtype=torch.uint8, device="cuda")
    t0 = time.time()
    done = 0
    while done < periods:
        n = min(P, periods - done)
        for k in range(n):
            b.gen_commands_into(stage.data_ptr() + k * W * R * 32, done + k, kind)
        b.step_sequence_device(stage.data_ptr(), n, 30)
        done += n
    st = b.reduce_stats()
    dt = time.time() - t0
    print("W=%d R=%d kind=%d: %d world-steps each in %.2f s (%.3e world-steps/s incl. generators); stats %s" %
          (W, R, kind, periods * 30, dt, W * periods * 30 / dt, {k: v for k, v in st.items() if k != "checksum"}), flush=True)
    assert st["bad_worlds"] == 0
    b.close()
run(4096, 12, 2000, 0)          # 60,000 world-steps = 1,000 s of play per world
run(4096, 12, 1000, 1)          # kicks and dribbling
run(4096, 22, 300, 1, "FIELD_DIV_A", 1)
run(1048576, 12, 10, 0)         # configs[4] on one GPU: 1 Mi worlds
