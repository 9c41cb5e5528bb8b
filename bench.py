#!/usr/bin/env python
"""bench.py -- world-steps/sec of the batched SSL world-stepper (BASELINE.json metric), strong scaling.

Workload = BASELINE.json configs[4]: 1,048,576 6v6 worlds on FIELD_2015 in total, sharded
over the N GPUs in contiguous ranges (N = 1: all of them on one GPU), randomized wheel-velocity commands redrawn every
30 world-steps, no per-step collective, one all-gather of the episode statistics at the end.
One bench *step* = 10 command periods x 30 world-steps = 300 world-steps of every world.  One world-step = one
`world_step` of the reference (src/sslsim.cpp:307-310: 1/60 s, two 1/120 s substeps).

  value     device-resident arm: the 10 command blocks of a step are generated into HBM (outside the event pair), then
            ONE step_worlds launch runs the 300 world-steps (world_batch_step_sequence); CUDA events bracket each launch on
            the batch's own stream; max over ranks of the summed launch times.  The working set of a launch (state +
            10 command blocks, >= 0.5 GB per GPU even at N = 8) is larger than L2, so no flush is needed.
  e2e       the same work through the C ABI with HOST buffers, closed loop per command period: H2D of the period's commands
            from pinned host memory, 30 world-steps, D2H of the per-world event / statistics words -- `e2e` -- or of those
            plus every pose and velocity (832 B per world, what the reference's loop reads after each step,
            src/main.c:80-85) -- `e2e.with_state`.  Chunked pipeline: `--e2e-chunks` world ranges per GPU, one stream each.
  parity    64 sampled worlds of the TIMED batch (global ids k * total/64) are re-run on the CPU oracle through the
            same schedule and compared bit for bit with the device state after the device arm and after the e2e arm.
  roofline  algorithmic bytes (1,232 B per 6v6 world-step) / launch duration vs measured HBM bandwidth.
  cpu_baseline  the CPU oracle (restated-Bullet port; Bullet itself is not buildable here) on the host cores for a
            bounded sample of the same workload (N = 1, rank 0 only).
  configs_1 / config_3 / config_4  sub-results at N = 1: BASELINE.json configs[1], [2], [3] at their stated sizes.
  cold_start  sub-result at N = 1: the headline workload with warmstart_factor = 0 (the cold-start step of spec v2 that
            rounds 1 and 2 measured); everything else runs the default parameters, i.e. Bullet's warm starting (spec v3).

`--impl reference` times that CPU oracle alone (rank 0 only) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PERIOD_STEPS = 30              # commands are redrawn every 30 world-steps
PERIODS_PER_STEP = 10          # one bench step = 10 command periods, fused into one launch by the device arm
WORLD_STEPS_PER_STEP = PERIOD_STEPS * PERIODS_PER_STEP
SETTLE_STEPS = 120             # drop from 0.5 m and settle before timing (2 s of sim time, command period 0)
TOTAL_WORLDS = 1 << 20         # 1,048,576 (BASELINE.json configs[4])
SEED = 0x5351
N_SAMPLED = 64


def bytes_per_world_step(R):
    """state read + write (32 B per body each way), command read (32 B per robot), aux RMW."""
    return 2 * 32 * (R + 1) + 32 * R + 16


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(R, sample_worlds, target_seconds, threads):
    """world-steps/s of the CPU oracle on a bounded sample of the workload (cpu_baseline leg)."""
    import oracle_binding as ob
    ob.build()
    w = ob.Worlds(sample_worlds, R, seed=SEED)
    w.gen_commands(0, 0)
    w.step(SETTLE_STEPS, threads=threads)
    period, done, t0 = 1, 0, time.perf_counter()
    while True:
        w.gen_commands(period, 0)
        w.step(PERIOD_STEPS, threads=threads)
        period += 1
        done += sample_worlds * PERIOD_STEPS
        el = time.perf_counter() - t0
        if el >= target_seconds:
            return done / el, el, period - 1


def config_dict(args, world_size):
    per = args.total_worlds // world_size
    return {"workload": "configs[4]: %d batched 6v6 worlds in total over %d GPU(s) (%d per GPU, contiguous ranges), "
                        "FIELD_2015, randomized wheel-velocity commands redrawn every %d world-steps; 1 bench step = "
                        "%d command periods = %d world-steps of every world"
                        % (args.total_worlds, world_size, per, PERIOD_STEPS, PERIODS_PER_STEP, WORLD_STEPS_PER_STEP),
            "total_worlds": args.total_worlds, "worlds_per_gpu": per, "robots_per_world": args.robots,
            "world_steps_per_step": WORLD_STEPS_PER_STEP, "command_periods_per_step": PERIODS_PER_STEP,
            "substeps_per_world_step": 2,
            "step": "docs/STEP_SPEC.md v3, default parameters: Bullet's warm starting on (warmstart_factor 0.85)",
            "sharding": "contiguous world ranges, no per-step collective; NCCL all-gather of the episode statistics",
            "l2": "no flush needed: one launch streams %.0f MiB per GPU (state + %d command blocks) > 126 MB L2"
                  % (per * (2 * 16 * (args.robots + 1) + 32 + PERIODS_PER_STEP * 32 * args.robots) / 2 ** 20,
                     PERIODS_PER_STEP)}


def run_reference(args, rank, world_size):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    Bullet (the reference's step, src/sslsim.cpp:323) cannot be built here (probe of the GPU box:
    profiles/r2_bullet_probe.txt), so this is the restated-Bullet CPU oracle (kind "port") with every host thread.
    Each step is a bounded sample: `--ref-sample-worlds` of the workload's worlds run the step's 300 world-steps;
    `value` is the measured rate, `ms_per_step` is rescaled to the full workload."""
    if rank != 0:
        return
    import oracle_binding as ob
    ob.build()
    threads = os.cpu_count() or 1
    R = args.robots
    sample = min(args.total_worlds, args.ref_sample_worlds)
    ids = [k * (args.total_worlds // sample) for k in range(sample)]
    w = ob.Worlds(sample, R, seed=SEED, world_ids=ids)
    w.gen_commands(0, 0)
    w.step(SETTLE_STEPS, threads=threads)

    def one_step(k):
        for j in range(PERIODS_PER_STEP):
            w.gen_commands(1 + k * PERIODS_PER_STEP + j, 0)
            w.step(PERIOD_STEPS, threads=threads)
    for k in range(args.warmup):
        one_step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(args.warmup + k)
    el = time.perf_counter() - t0
    value = sample * WORLD_STEPS_PER_STEP * args.steps / el
    cfg = config_dict(args, world_size)
    cfg["reference_sample"] = ("the CPU arm steps %d of the %d worlds (ids k * %d) per bench step and rescales "
                               "ms_per_step by %d; value is the measured rate" %
                               (sample, args.total_worlds, args.total_worlds // sample, args.total_worlds // sample))
    line = {
        "impl": "reference", "metric": "world-steps/sec (6v6, 60 Hz)", "value": value, "unit": "world-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el * 1e3 / args.steps * (args.total_worlds / sample), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d of %d worlds x %d world-steps per step, %d steps in %.1f s; restated-Bullet CPU "
                                   "oracle (oracle/ssl_oracle.c, gcc -O2, pthreads), Bullet itself not buildable here" %
                                   (sample, args.total_worlds, WORLD_STEPS_PER_STEP, args.steps, el)},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--total-worlds", type=int, default=TOTAL_WORLDS, help="worlds in total over all GPUs (configs[4]: 1,048,576)")
    ap.add_argument("--robots", type=int, default=12, help="robots per world (6v6 = 12)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per world: 16, 32 or 0 = auto")
    ap.add_argument("--ref-sample-worlds", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sub-results", action="store_true", help="skip the configs[1], [2], [3] sub-results (N = 1 only anyway)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="e2e arm: chunks of worlds per GPU, each on its own stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world_size)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    sim = importlib.import_module("ssl-sim_b200")
    if not os.path.exists(sim.LIB_PATH):
        sim.build()   # fresh checkout: compile the CUDA library (nvcc); there is still no fallback path
    sim.lib()  # fails loudly if the CUDA library is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world_size > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    R, K, WU, PPS = args.robots, args.steps, args.warmup, PERIODS_PER_STEP
    TOTAL = args.total_worlds
    bpws = bytes_per_world_step(R)
    first_world, W = sim.shard_range(TOTAL, rank, world_size)   # this GPU's contiguous range
    L = sim.lib()
    # one process per GPU through the C ABI's multi-GPU layer (include/sslsim_multi.h): rank 0's ncclUniqueId reaches the
    # other ranks over torch.distributed (plumbing); the shard, its step and the ncclAllGather of the episode statistics
    # are the library's own
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(sim.WorldMulti.unique_id()), dtype=torch.uint8))
    if distributed:
        dist.broadcast(uid, 0)
    multi = sim.WorldMulti(TOTAL, R, rank=rank, n_ranks=world_size, device=local_rank, nccl_id=bytes(uid.cpu().numpy()), seed=SEED)
    b = multi.shard(0, seed=SEED)
    assert (b.W, b.first_world_id) == (W, first_world)
    if args.lanes:
        b.set_lanes_per_world(args.lanes)
    cmd_bytes = W * R * 32
    stage = torch.empty(PPS * cmd_bytes, dtype=torch.uint8, device=dev)   # the command blocks of one bench step

    def gen_step(k):   # device-side synthetic command stream: periods 1 + k*PPS ... of the keyed generator
        for j in range(PPS):
            b.gen_commands_into(stage.data_ptr() + j * cmd_bytes, 1 + k * PPS + j, 0)

    def barrier():
        b.sync()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def allgather_stats():
        _, tot = multi.allgather_stats()     # K4 on the device + ncclAllGather inside the library
        return tot

    def allmax(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # ---- sampled worlds for the in-run parity check (global ids k * TOTAL/64; this rank checks the ones it owns)
    stride = max(1, TOTAL // N_SAMPLED)
    sampled = [g for g in range(0, TOTAL, stride)][:N_SAMPLED]
    mine = [g for g in sampled if first_world <= g < first_world + W]
    local_ids = [g - first_world for g in mine]

    # ---- settle + warm-up (device arm) ---------------------------------------------------------------------
    b.gen_commands(0, 0)
    b.step(SETTLE_STEPS)
    for k in range(WU):
        gen_step(k)
        b.step_sequence_device(stage.data_ptr(), PPS, PERIOD_STEPS)
    allgather_stats()  # warms the NCCL communicator up
    snap = b.snapshot()   # state after settle + warm-up: the e2e arms start here too

    # ---- device-resident arm -----------------------------------------------------------------------------
    b.enable_kernel_timing(True)
    b.kernel_time(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for k in range(K):
        gen_step(WU + k)                                                   # untimed: outside the event pair
        b.step_sequence_device(stage.data_ptr(), PPS, PERIOD_STEPS)        # timed: CUDA events around the launch
    total_stats = allgather_stats()
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.finish()
    kern_ms, launches = b.kernel_time(reset=True)
    b.enable_kernel_timing(False)
    kern_ms_max, wall_ms_max = allmax(kern_ms, wall_s * 1e3)
    per_rank_ms = [kern_ms]
    if distributed:
        tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world_size)]
        dist.all_gather(tl, torch.tensor([kern_ms], dtype=torch.float64, device=dev))
        per_rank_ms = [float(t[0]) for t in tl]
    units = TOTAL * WORLD_STEPS_PER_STEP * K
    value = units / (kern_ms_max * 1e-3)
    dev_state = b.read_worlds(local_ids) if local_ids else None

    # ---- end-to-end arms: host buffers through the C ABI ------------------------------------------------------------
    e2e, e2e_state = None, None
    N_HOST = 2
    if not args.no_e2e:
        nch = max(1, min(args.e2e_chunks, W))
        host_cmd = torch.empty(N_HOST * cmd_bytes, dtype=torch.uint8).pin_memory()
        for s in range(N_HOST):   # host command sets = generator periods 100000 + s (copied out once, untimed)
            b.gen_commands_into(stage.data_ptr(), 100000 + s, 0)
            b.sync()
            host_cmd[s * cmd_bytes:(s + 1) * cmd_bytes].copy_(stage[:cmd_bytes])
        torch.cuda.synchronize()
        host_aux = torch.empty(W * 8, dtype=torch.int32).pin_memory()
        host_pos = torch.empty(W * (R + 1) * 4, dtype=torch.float32).pin_memory()
        host_vel = torch.empty(W * (R + 1) * 4, dtype=torch.float32).pin_memory()
        hp, ap_, pp, vp = host_cmd.data_ptr(), host_aux.data_ptr(), host_pos.data_ptr(), host_vel.data_ptr()
        h = b._h
        NB = (R + 1) * 16

        def run_e2e(with_state):
            b.restore(snap)
            b.set_chunks(nch)
            ranges = [b.chunk_range(c) for c in range(nch)]
            counter = [0]

            def period():
                base = hp + (counter[0] % N_HOST) * cmd_bytes
                counter[0] += 1
                for c, (first, count) in enumerate(ranges):
                    # closed loop: this chunk's results of the previous period are in host memory before its next
                    # commands are submitted; the other chunks' queued work overlaps the wait
                    rc = L.world_batch_chunk_wait(h, c)
                    if rc >= 0:
                        rc = L.world_batch_chunk_step_ex(
                            h, c, C.c_void_p(base + first * R * 32), PERIOD_STEPS, C.c_void_p(ap_ + first * 32),
                            C.c_void_p(pp + first * NB) if with_state else None,
                            C.c_void_p(vp + first * NB) if with_state else None)
                    if rc < 0:
                        raise RuntimeError("e2e chunk step failed: %d" % rc)
            for _ in range(WU * PPS):
                period()
            barrier()
            t1 = time.perf_counter()
            for _ in range(K * PPS):
                period()
            barrier()                      # every chunk's last results are in host memory
            el = allmax(time.perf_counter() - t1)[0]
            st = b.reduce_stats()
            state = b.read_worlds(local_ids) if local_ids else None
            b.set_chunks(0)
            return el, st, state

        e_s, e_stats, e2e_state = run_e2e(False)
        s_s, s_stats, _ = run_e2e(True)
        assert s_stats == e_stats, "the two e2e variants must produce identical worlds"
        api = ("%d chunks per GPU x [world_batch_chunk_wait + world_batch_chunk_step_ex(host commands, %d steps, host aux%s)], "
               "one CUDA stream per chunk, per command period")
        e2e = {"value": units / e_s, "unit": "world-steps/s",
               "h2d_bytes_per_step": PPS * cmd_bytes * world_size, "d2h_bytes_per_step": PPS * W * 32 * world_size,
               "ms_per_step": e_s * 1e3 / K, "chunks": nch, "api": api % (nch, PERIOD_STEPS, ""),
               "commands": "%d host command sets (generator periods 100000..) used in turn -- a full step's worth of distinct "
                           "host blocks would need %.0f GB of pinned memory -- so the e2e worlds follow a different command "
                           "stream than the device arm's (fresh draws every period) and cost a few percent less per step; the "
                           "two arms are checked against the oracle separately (parity)" % (N_HOST, K * PPS * cmd_bytes / 1e9),
               "with_state": {"value": units / s_s, "unit": "world-steps/s", "ms_per_step": s_s * 1e3 / K,
                              "h2d_bytes_per_step": PPS * cmd_bytes * world_size,
                              "d2h_bytes_per_step": PPS * W * (32 + 2 * NB) * world_size,
                              "api": api % (nch, PERIOD_STEPS, ", host pos, host vel"),
                              "note": "every pose and velocity read back each command period (832 B per world), as the "
                                      "reference's loop does after each step (src/main.c:80-85)"}}
        del host_cmd, host_aux, host_pos, host_vel
    b.free_snapshot(snap)

    # ---- in-run parity: the sampled worlds of the timed batch on the CPU oracle ---------------------------------------
    parity = None
    if not args.no_parity:
        import oracle_binding as ob
        ob.build()
        threads = os.cpu_count() or 1
        bad_dev = bad_e2e = skipped = 0
        if mine:
            o = ob.Worlds(len(mine), R, seed=SEED, world_ids=mine)
            o.gen_commands(0, 0)
            o.step(SETTLE_STEPS, threads=threads)
            for p in range(WU * PPS):
                o.gen_commands(1 + p, 0)
                o.step(PERIOD_STEPS, threads=threads)
            at_snap = (o.pos.copy(), o.vel.copy(), o.aux.copy(), o.warm.copy())   # the whole state, warm tables included
            for p in range(WU * PPS, (WU + K) * PPS):
                o.gen_commands(1 + p, 0)
                o.step(PERIOD_STEPS, threads=threads)
            ok_rows = ((o.aux[:, 0] >> 16) & 1) == 0          # overflowed worlds: state unspecified (STEP_SPEC 4.3)
            skipped = int((~ok_rows).sum())
            same = [all(np.array_equal(a[i], g[i]) for a, g in zip((o.pos, o.vel, o.aux), dev_state)) for i in range(len(mine))]
            bad_dev = sum(1 for i, s in enumerate(same) if ok_rows[i] and not s)
            if e2e_state is not None:
                o.pos[:], o.vel[:], o.aux[:], o.warm[:] = at_snap
                for p in range((WU + K) * PPS):
                    o.gen_commands(100000 + p % N_HOST, 0)
                    o.step(PERIOD_STEPS, threads=threads)
                same = [all(np.array_equal(a[i], g[i]) for a, g in zip((o.pos, o.vel, o.aux), e2e_state)) for i in range(len(mine))]
                ok2 = ((o.aux[:, 0] >> 16) & 1) == 0
                bad_e2e = sum(1 for i, s in enumerate(same) if ok2[i] and not s)
        t = torch.tensor([bad_dev, bad_e2e, skipped, len(mine)], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bad_dev, bad_e2e, skipped, n_checked = (int(v) for v in t)
        parity = {"worlds": n_checked, "bit_exact": bad_dev == 0 and bad_e2e == 0, "mismatching_device_arm": bad_dev,
                  "mismatching_e2e_arm": bad_e2e if e2e is not None else None, "overflowed_skipped": skipped,
                  "world_steps_checked_per_world": SETTLE_STEPS + (WU + K) * WORLD_STEPS_PER_STEP,
                  "against": "oracle/ssl_oracle.c (restated Bullet 2.82 step, parity unpinned against Bullet itself)",
                  "sampled_ids": "k * %d, k = 0..%d" % (stride, len(sampled) - 1)}

    # ---- sub-results and CPU baseline (N = 1, rank 0) ---------------------------------------------------------------------
    nccl_version = L.sslsim_multi_nccl_version().decode()
    b.close()
    multi.close()
    del stage
    subs = {}
    cpu = None
    if rank == 0 and world_size == 1:
        if not args.no_sub_results:
            import bench_configs
            subs = bench_configs.run_all(sim, dev, local_rank, SEED, check_parity=not args.no_parity)
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sw = min(TOTAL, args.ref_sample_worlds)
            v, el, reps = cpu_oracle_rate(R, sw, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": "world-steps/s", "cores": threads, "kind": "port",
                   "sample": "%d worlds x %d world-steps x %d command periods in %.1f s; restated-Bullet CPU oracle "
                             "(oracle/ssl_oracle.c, gcc -O2, pthreads); Bullet itself is not buildable here (SURVEY 8c)"
                             % (sw, PERIOD_STEPS, reps, el)}

    if rank == 0:
        peak, peak_src = measured_peak()
        slow = max(range(world_size), key=lambda r: per_rank_ms[r])
        per_launch_bytes = bpws * W * WORLD_STEPS_PER_STEP
        per_launch_s = kern_ms_max * 1e-3 / launches
        achieved = per_launch_bytes / per_launch_s / 1e9
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    tj = json.load(f)
                wsl = float(tj["world_steps_per_launch"])
                traffic = tj["dram_bytes_per_launch"] * (W * WORLD_STEPS_PER_STEP) / wsl   # scaled to this launch shape
                ipw = tj["warp_instructions_per_launch"] / wsl
                peak_issue = 148 * 4 * (clocks.get("sm_max_mhz") or 1965) * 1e6
                issue = {"warp_instructions_per_world_step": ipw, "achieved_per_s": ipw * value / world_size,
                         "peak_per_s": peak_issue, "frac": ipw * value / world_size / peak_issue,
                         "source": "profiles/traffic.json (ncu smsp__inst_executed.sum of %s)" % tj.get("capture", "the committed capture")}
            except Exception:
                traffic = None
        line = {
            "metric": "world-steps/sec (6v6, 60 Hz)", "value": value, "unit": "world-steps/s",
            "n_gpus": world_size, "steps": K, "warmup": WU, "ms_per_step": kern_ms_max / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, world_size),
            "timed_region_s": kern_ms_max * 1e-3, "wall_ms_per_step_incl_command_generation": wall_ms_max / K,
            "clocks": clocks, "e2e": e2e, "parity": parity,
            "gpu_launches": launches + K * PPS + 1,
            "gpu_launches_detail": {"step_worlds": launches, "gen_commands": K * PPS, "reduce_stats": 1},
            "per_rank_kernel_ms": {"min": min(per_rank_ms), "median": statistics.median(per_rank_ms),
                                   "max": max(per_rank_ms), "slowest_rank": slow, "all": per_rank_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "step_worlds", "algorithmic_bytes_per_launch": per_launch_bytes,
                         "bytes_per_world_step": bpws, "world_steps_per_launch": W * WORLD_STEPS_PER_STEP,
                         "kernel_ms_per_launch": kern_ms_max / launches, "issue_ceiling": issue,
                         "note": "latency/issue-bound fused kernel: state stays in registers for the 300 world-steps of "
                                 "a launch (DRAM traffic = state once + one command block per period); issue_ceiling "
                                 "is the informative bound; see DESIGN.md"},
            "cpu_baseline": cpu,
            "episode_stats": total_stats,
            "multi_gpu": {"api": "include/sslsim_multi.h: new_world_multi_rank + world_multi_allgather_stats (ncclAllGather of "
                                 "8 x u64 per GPU, once per arm), no per-step collective", "nccl": nccl_version},
        }
        line.update(subs)
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
