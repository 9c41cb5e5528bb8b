"""Sub-results of bench.py at N = 1: BASELINE.json configs[1], [2], [3] at their stated
sizes, each with an in-run oracle check of sampled worlds of the timed batch.  Imported by bench.py only."""
import ctypes as C
import os
import time

import numpy as np
import torch

PERIOD_STEPS = 30


def _parity(sim, b, ids, R, seed, schedule, fld=None, mode=0, params_kw=None, kind=0):
    """Re-run worlds `ids` of batch b on the CPU oracle through `schedule` = [(period index, steps), ...] and compare
    bit for bit with the device state."""
    import oracle_binding as ob
    ob.build()
    threads = os.cpu_count() or 1
    p = ob.default_params(**(params_kw or {}))
    o = ob.Worlds(len(ids), R, fld=getattr(ob, fld) if fld else ob.FIELD_2015, params=p, seed=seed, mode=mode, world_ids=ids)
    for period, steps in schedule:
        o.gen_commands(period, kind)
        o.step(steps, threads=threads)
    pos, vel, aux = b.read_worlds(ids)
    ok = ((o.aux[:, 0] >> 16) & 1) == 0
    bad = sum(1 for i in range(len(ids)) if ok[i] and not (np.array_equal(o.pos[i], pos[i]) and
                                                          np.array_equal(o.vel[i], vel[i]) and np.array_equal(o.aux[i], aux[i])))
    return {"worlds": len(ids), "bit_exact": bad == 0, "mismatching": bad, "overflowed_skipped": int((~ok).sum()),
            "world_steps_checked_per_world": sum(s for _, s in schedule)}, o


def config_1(sim, dev, device, seed, check_parity, K=40, PPL=10, WU=3, n_chunks=4):
    """configs[1]: 4,096 batched 6v6 worlds on one GPU, randomized wheel commands, parity on ids = 0 mod 64."""
    W, R = 4096, 12
    b = sim.WorldBatch(W, R, device=device, seed=seed)
    cmd_bytes = W * R * 32
    n_sets = (WU + K) * PPL
    stage = torch.empty(n_sets * cmd_bytes, dtype=torch.uint8, device=dev)
    for k in range(n_sets):
        b.gen_commands_into(stage.data_ptr() + k * cmd_bytes, 1 + k, 0)
    b.gen_commands(0, 0)
    b.step(120)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)   # working set 4.9 MiB < L2: flush between launches
    snap0 = b.snapshot()

    def fused(periods_per_launch):
        b.restore(snap0)
        for k in range(WU):
            b.step_sequence_device(stage.data_ptr() + k * PPL * cmd_bytes, PPL, PERIOD_STEPS)
        b.enable_kernel_timing(True)
        b.kernel_time(reset=True)
        b.sync()
        p = WU * PPL
        while p < n_sets:
            b.step_sequence_device(stage.data_ptr() + p * cmd_bytes, periods_per_launch, PERIOD_STEPS)
            b.sync()
            flush.zero_()
            torch.cuda.synchronize()
            p += periods_per_launch
        ms, n = b.kernel_time(reset=True)
        b.enable_kernel_timing(False)
        return W * PERIOD_STEPS * K * PPL / (ms * 1e-3), ms / n, n, b.reduce_stats()

    v_fused, ms_launch, n_launch, st_fused = fused(PPL)
    par = None
    if check_parity:
        ids = list(range(0, W, 64))
        par, _ = _parity(sim, b, ids, R, seed, [(0, 120)] + [(1 + k, PERIOD_STEPS) for k in range(n_sets)])
    v_single, _, _, st_single = fused(1)
    assert st_single == st_fused, "fused and per-period launches must produce identical worlds"

    # staggered chunks: each chunk runs its own 10-period sequences back to back on its own stream, so the slow
    # worlds at the end of one chunk's launch overlap the other chunks' work (no batch-wide barrier per launch)
    b.restore(snap0)
    b.set_chunks(n_chunks)
    for k in range(WU):
        for c in range(n_chunks):
            b.chunk_step_sequence(c, stage.data_ptr() + k * PPL * cmd_bytes, PPL, PERIOD_STEPS)
    b.sync()
    flush.zero_()
    torch.cuda.synchronize()
    b.timer_start()
    for k in range(WU, WU + K):
        for c in range(n_chunks):
            b.chunk_step_sequence(c, stage.data_ptr() + k * PPL * cmd_bytes, PPL, PERIOD_STEPS)
    b.sync()
    ms_ch = b.timer_stop()
    st_chunks = b.reduce_stats()
    b.set_chunks(0)
    assert st_chunks == st_fused, "chunked sequences must produce identical worlds"
    v_chunks = W * PERIOD_STEPS * K * PPL / (ms_ch * 1e-3)

    # closed loop through host buffers, one command period at a time (8 chunks)
    L, h = sim.lib(), b._h
    n_host = 16
    host_cmd = torch.empty(n_host * cmd_bytes, dtype=torch.uint8).pin_memory()
    host_cmd.copy_(stage[: n_host * cmd_bytes].cpu())
    host_aux = torch.empty(W * 8, dtype=torch.int32).pin_memory()
    hp, ap_ = host_cmd.data_ptr(), host_aux.data_ptr()
    b.restore(snap0)
    b.set_chunks(8)
    ranges = [b.chunk_range(c) for c in range(8)]

    def period(k):
        base = hp + (k % n_host) * cmd_bytes
        for c, (first, count) in enumerate(ranges):
            if L.world_batch_chunk_wait(h, c) < 0 or L.world_batch_chunk_step(
                    h, c, C.c_void_p(base + first * R * 32), PERIOD_STEPS, C.c_void_p(ap_ + first * 32)) < 0:
                raise RuntimeError("e2e chunk step failed")
    for k in range(30):
        period(k)
    b.sync()
    t0 = time.perf_counter()
    for k in range(30, 30 + K * PPL):
        period(k)
    b.sync()
    e2e_s = time.perf_counter() - t0
    b.set_chunks(0)
    b.free_snapshot(snap0)
    b.close()
    return {"workload": "configs[1]: 4,096 batched 6v6 worlds on 1 GPU, wheel commands redrawn every 30 world-steps; "
                        "%d command periods timed; L2 flushed between launches (working set 4.9 MiB)" % (K * PPL),
            "value": v_chunks, "unit": "world-steps/s",
            "how": "%d staggered chunks (world_batch_chunk_step_sequence, 10 periods per launch, one stream per chunk), "
                   "CUDA events around the whole region" % n_chunks,
            "one_stream_fused_10_periods": {"value": v_fused, "kernel_ms_per_launch": ms_launch, "launches": n_launch},
            "one_launch_per_period": {"value": v_single},
            "e2e": {"value": W * PERIOD_STEPS * K * PPL / e2e_s, "unit": "world-steps/s",
                    "h2d_bytes_per_period": cmd_bytes, "d2h_bytes_per_period": W * 32,
                    "api": "8 chunks x [world_batch_chunk_wait + world_batch_chunk_step(host commands, 30, host aux)]"},
            "parity": par}


def config_3(sim, dev, device, seed, check_parity, K=3, PPL=10):
    """configs[2]: 65,536 batched 11v11 Division-A worlds, crowded-penalty-area starts, kicks + dribbling."""
    W, R = 65536, 22
    b = sim.WorldBatch(W, R, fld="FIELD_DIV_A", device=device, seed=seed, init_mode=1)
    cmd_bytes = W * R * 32
    stage = torch.empty(PPL * cmd_bytes, dtype=torch.uint8, device=dev)
    sched = []

    def step(k):
        for j in range(PPL):
            b.gen_commands_into(stage.data_ptr() + j * cmd_bytes, k * PPL + j, 1)
            sched.append((k * PPL + j, PERIOD_STEPS))
        b.step_sequence_device(stage.data_ptr(), PPL, PERIOD_STEPS)
    step(0)                                   # warm-up
    b.enable_kernel_timing(True)
    b.kernel_time(reset=True)
    for k in range(1, 1 + K):
        step(k)
    ms, n = b.kernel_time(reset=True)
    st = b.reduce_stats()
    par = None
    if check_parity:
        ids = list(range(0, W, W // 16))
        par, _ = _parity(sim, b, ids, R, seed, sched, fld="FIELD_DIV_A", mode=1, kind=1)
    b.close()
    bpws = 2 * 32 * (R + 1) + 32 * R + 16
    v = W * PERIOD_STEPS * PPL * K / (ms * 1e-3)
    return {"workload": "configs[2]: 65,536 batched 11v11 worlds on FIELD_DIV_A, all bodies starting inside one penalty "
                        "area, wheel + kick + dribble commands; %d launches of 10 command periods" % K,
            "value": v, "unit": "world-steps/s", "kernel_ms_per_launch": ms / n, "bytes_per_world_step": bpws,
            "contacts_per_substep_per_world": st["contacts"] / (W * 2.0 * PERIOD_STEPS * PPL * (K + 1)),
            "bad_worlds": st["bad_worlds"], "parity": par}


def config_4(sim, dev, device, seed, check_parity, n_steps=600):
    """configs[3]: 16,384 6v6 worlds, per-step kicker + dribbler activation, detection gather of 64 observed worlds
    (ids k * 256) into host memory after EVERY world-step (closed loop: one launch + gather + D2H per world-step)."""
    W, R = 16384, 12
    b = sim.WorldBatch(W, R, device=device, seed=seed)
    observed = [k * 256 for k in range(64)]
    sched = []
    for k in range(4):                        # settle + warm-up
        b.gen_commands(k, 1)
        b.step(PERIOD_STEPS)
        sched.append((k, PERIOD_STEPS))
    b.gather_detections(observed)
    b.sync()
    t0 = time.perf_counter()
    det = None
    for s in range(n_steps):
        if s % PERIOD_STEPS == 0:
            b.gen_commands(4 + s // PERIOD_STEPS, 1)
            sched.append((4 + s // PERIOD_STEPS, PERIOD_STEPS))
        b.step(1)
        det = b.gather_detections(observed)
    b.sync()
    el = time.perf_counter() - t0
    st = b.reduce_stats()
    par = None
    if check_parity:
        ids = observed[::4]
        par, o = _parity(sim, b, ids, R, seed, sched, kind=1)
        rid = [i // 2 for i in range(R)]
        par["gather_matches_oracle"] = all(np.array_equal(det[4 * i], o.gather(i, rid)) for i in range(len(ids)))
    b.close()
    return {"workload": "configs[3]: 16,384 batched 6v6 worlds, spinner on + random kicks every period, detection records of "
                        "64 observed worlds gathered to host memory after every world-step; %d world-steps" % n_steps,
            "value": W * n_steps / el, "unit": "world-steps/s", "us_per_world_step_launch_gather_d2h": el * 1e6 / n_steps,
            "kicks": st["kicks"], "goals": st["goals_blue"] + st["goals_yellow"], "ball_out": st["ball_out"],
            "gather_bytes_per_step": 64 * (6 + 7 * R) * 4, "parity": par}


def cold_start(sim, dev, device, seed, check_parity, K=3, PPL=10, W=1 << 20):
    """The headline workload (configs[4]: 1,048,576 6v6 worlds, wheel commands) with warmstart_factor = 0, i.e. the
    cold-start step of spec v2 that rounds 1 and 2 measured: what Bullet's warm starting (spec v3, the default) costs."""
    R = 12
    b = sim.WorldBatch(W, R, device=device, seed=seed, params=sim.default_params(warmstart_factor=0.0))
    cmd_bytes = W * R * 32
    stage = torch.empty(PPL * cmd_bytes, dtype=torch.uint8, device=dev)
    sched = [(0, 120)]
    b.gen_commands(0, 0)
    b.step(120)

    def step(k):
        for j in range(PPL):
            b.gen_commands_into(stage.data_ptr() + j * cmd_bytes, 1 + k * PPL + j, 0)
            sched.append((1 + k * PPL + j, PERIOD_STEPS))
        b.step_sequence_device(stage.data_ptr(), PPL, PERIOD_STEPS)
    step(0)                                   # warm-up
    b.enable_kernel_timing(True)
    b.kernel_time(reset=True)
    for k in range(1, 1 + K):
        step(k)
    ms, n = b.kernel_time(reset=True)
    par = None
    if check_parity:
        par, _ = _parity(sim, b, list(range(0, W, W // 16)), R, seed, sched, params_kw={"warmstart_factor": 0.0})
    b.close()
    return {"workload": "configs[4] with SslSimParams.warmstart_factor = 0 (every contact row starts from zero impulse: the "
                        "step of spec v2, rounds 1-2): %d 6v6 worlds, %d launches of 10 command periods" % (W, K),
            "value": W * PERIOD_STEPS * PPL * K / (ms * 1e-3), "unit": "world-steps/s", "kernel_ms_per_launch": ms / n,
            "parity": par}


def run_all(sim, dev, device, seed, check_parity=True):
    out = {}
    for name, fn in (("configs_1", config_1), ("config_3", config_3), ("config_4", config_4), ("cold_start", cold_start)):
        try:
            out[name] = fn(sim, dev, device, seed, check_parity)
        except Exception as e:   # a sub-result must not take the headline line down
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out
