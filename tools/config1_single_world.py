"""configs[0]: a single 6v6 world on FIELD_2015 stepped 10,000 times at 60 Hz through the reference's own API
(new_world / world_add_robot / world_step, reference src/main.c:62-80), and the same 10,000 steps as one launch."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
L = sim.lib()
w = L.new_world(C.byref(sim.field("FIELD_2015")))
for i in range(6):
    L.world_add_robot(w, i, 0); L.world_add_robot(w, i, 1)
for _ in range(200):
    L.world_step(w)
L.world_get_frame_number(w)
t0 = time.perf_counter()
for _ in range(10000):
    L.world_step(w)                       # one launch each, asynchronous
b = L.ball_get_vec(L.world_get_game_ball(w))   # a getter synchronises
dt = time.perf_counter() - t0
print("legacy API, 10,000 x world_step + one getter: %.1f ms -> %.3e world-steps/s (frame %d, ball z %.4f)" %
      (dt * 1e3, 10000 / dt, L.world_get_frame_number(w), b.z))
t0 = time.perf_counter()
for _ in range(1000):
    L.world_step(w)
    L.ball_get_vec(L.world_get_game_ball(w))   # closed loop: read the state back after every step
dt = time.perf_counter() - t0
print("legacy API, closed loop (world_step + ball_get_vec) x 1,000: %.1f us per step -> %.3e world-steps/s" % (dt * 1e3, 1000 / dt))
bb = sim.WorldBatch(1, 12)
bb.step(200); bb.sync()
bb.timer_start(); bb.step(10000); ms = bb.timer_stop()
print("batch of 1, world_batch_step(10000) in one launch: %.1f ms -> %.3e world-steps/s" % (ms, 10000 / (ms * 1e-3)))
