"""Scratch timing of the step kernel at a few batch sizes / lane mappings (not the bench)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")

def run(W, R, lanes, steps=30, reps=10, kind=0, fld="FIELD_2015", mode=0, settle=120):
    b = sim.WorldBatch(W, R, fld=fld, init_mode=mode)
    b.set_lanes_per_world(lanes)
    b.gen_commands(0, kind)
    b.step(settle)
    b.sync()
    best = 1e9
    tot = 0.0
    for r in range(reps):
        b.gen_commands(r + 1, kind)
        b.timer_start()
        b.step(steps)
        ms = b.timer_stop()
        best = min(best, ms); tot += ms
    ws = W * steps * reps / (tot * 1e-3)
    print("W=%7d R=%2d lanes=%2d kind=%d mode=%d: avg %.3f ms/launch (best %.3f) -> %.3e world-steps/s, contacts/substep %.2f"
          % (W, R, lanes, kind, mode, tot / reps, best, ws,
             b.reduce_stats()["contacts"] / (W * (settle + steps * reps) * 2.0)), flush=True)
    b.close()

if __name__ == "__main__":
    for W in (4096, 16384, 131072):
        for lanes in (16, 32):
            run(W, 12, lanes)
    run(16384, 12, 16, kind=1)
    run(65536, 22, 32, kind=1, fld="FIELD_DIV_A", mode=1)
