#!/usr/bin/env python
"""bench.py -- world-steps/sec of the batched SSL world-stepper (BASELINE.json metric).

Workload (BASELINE.json configs[1]): 4,096 batched 6v6 worlds per GPU on
FIELD_2015 with randomized wheel-velocity commands that are redrawn every 30 world-steps.
One bench *step* = one command batch + 30 world-steps of every world.  One world-step = one `world_step` of the reference
(reference src/sslsim.cpp:307-310: 1/60 s, two 1/120 s substeps).

  value     device-resident arm: the K command batches are staged in HBM before the timed region and
            consumed `--periods-per-launch` (default 10) at a time by one `step_worlds` launch
            (world_batch_step_sequence: every world runs through its periods without a batch-wide
            barrier); CUDA events bracket each launch on the batch's own stream; L2 is flushed between
            launches (outside the event pairs).  `one_launch_per_period` is the same K steps with one
            launch per command period -- what a closed-loop caller (and the e2e arm) does.
  e2e       the same work through the C ABI with HOST buffers, closed loop: per step the command batch is
            copied from pinned host memory, stepped, and the per-world event / statistics words are read
            back to host memory; wall time of K steps.  The batch is driven as `--e2e-chunks` (default 8)
            contiguous chunks of worlds, each on its own CUDA stream (world_batch_chunk_step /
            world_batch_chunk_wait): the host waits for a chunk's results of step k before it submits
            that chunk's step k+1, while the other chunks' copies and launches keep the GPU busy.
            `e2e.whole_batch` is the same loop with one stream and a batch-wide barrier per step
            (world_batch_set_commands + world_batch_step + world_batch_read_aux).
  roofline  algorithmic bytes (1,232 B per 6v6 world-step) / kernel duration vs the
            measured HBM copy bandwidth.
  cpu_baseline  the CPU oracle (restated-Bullet port; Bullet itself is not buildable here) on the
            host cores for a bounded sample of the same workload.

`--impl reference` times that CPU oracle alone (rank 0 only) on the same config.
Multi-GPU: one process per GPU (torchrun), contiguous world ranges, no per-step collective; one
NCCL all-gather of the episode statistics at the end of the timed region.
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

WORLD_STEPS_PER_STEP = 30      # commands are redrawn every 30 world-steps
SETTLE_STEPS = 120             # drop from 0.5 m and settle before timing (2 s of sim time)
SEED = 0x5351


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

    def _sample_once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if self.ok and not self.samples:   # a very short timed region: the GPU is still at its load clock right after it
            try:
                self._sample_once()
            except Exception:
                pass
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
        w.step(WORLD_STEPS_PER_STEP, threads=threads)
        period += 1
        done += sample_worlds * WORLD_STEPS_PER_STEP
        el = time.perf_counter() - t0
        if el >= target_seconds:
            return done / el, el, period - 1


def run_reference(args, rank, world_size):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    Bullet (the reference's step, src/sslsim.cpp:323) cannot be built here, so this is the
    restated-Bullet CPU oracle (kind "port") with every host thread, on the bench config."""
    if rank != 0:
        return
    import oracle_binding as ob
    ob.build()
    threads = os.cpu_count() or 1
    R = args.robots
    sample = min(args.worlds, args.ref_sample_worlds)
    w = ob.Worlds(sample, R, seed=SEED)
    w.gen_commands(0, 0)
    w.step(SETTLE_STEPS, threads=threads)
    for k in range(args.warmup):
        w.gen_commands(1 + k, 0)
        w.step(WORLD_STEPS_PER_STEP, threads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        w.gen_commands(1 + args.warmup + k, 0)
        w.step(WORLD_STEPS_PER_STEP, threads=threads)
    el = time.perf_counter() - t0
    value = sample * WORLD_STEPS_PER_STEP * args.steps / el
    line = {
        "impl": "reference", "metric": "world-steps/sec (6v6, 60 Hz)", "value": value, "unit": "world-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el * 1e3 / args.steps * (args.worlds / sample), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world_size),
        "cpu_baseline": {"value": value, "unit": "world-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d of %d worlds x %d world-steps per step, %d steps; restated-Bullet CPU oracle "
                                   "(oracle/ssl_oracle.c), Bullet itself not buildable here" %
                                   (sample, args.worlds, WORLD_STEPS_PER_STEP, args.steps)},
        "e2e": {"value": value, "unit": "world-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, world_size):
    return {"workload": "configs[1]: %d batched %dv%d worlds per GPU, FIELD_2015, randomized wheel-velocity "
                        "commands redrawn every %d world-steps; 1 bench step = %d world-steps of every world"
                        % (args.worlds, args.robots // 2, args.robots // 2, WORLD_STEPS_PER_STEP, WORLD_STEPS_PER_STEP),
            "worlds_per_gpu": args.worlds, "robots_per_world": args.robots,
            "world_steps_per_step": WORLD_STEPS_PER_STEP, "substeps_per_world_step": 2,
            "total_worlds": args.worlds * world_size, "sharding": "contiguous world ranges, no per-step collective",
            "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA-event pairs); "
                  "working set %.1f MiB < L2" % (args.worlds * (bytes_per_world_step(args.robots) - 16 + 32) / 2 ** 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds per GPU (configs[1]: 4096)")
    ap.add_argument("--robots", type=int, default=12, help="robots per world (6v6 = 12)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per world: 16, 32 or 0 = auto")
    ap.add_argument("--ref-sample-worlds", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8,
                    help="e2e arm: chunks of worlds, each on its own stream (1 = whole batch, one barrier per step)")
    ap.add_argument("--periods-per-launch", type=int, default=10,
                    help="device-resident arm: command periods (bench steps) fused into one step_worlds launch")
    ap.add_argument("--large-worlds", type=int, default=131072,
                    help="extra measurement at this many worlds per GPU (0 = skip); not the headline value")
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

    W, R, K, WU = args.worlds, args.robots, args.steps, args.warmup
    N = R + 1
    bpws = bytes_per_world_step(R)
    first_world, _ = sim.shard_range(W * world_size, rank, world_size)   # contiguous range per GPU
    b = sim.WorldBatch(W, R, device=local_rank, seed=SEED, first_world_id=first_world)
    if args.lanes:
        b.set_lanes_per_world(args.lanes)
    # stage every command batch (warm-up + timed) in HBM before the timed region
    cmd_bytes = W * R * 32
    n_sets = WU + K
    stage = torch.empty(n_sets * cmd_bytes, dtype=torch.uint8, device=dev)
    for k in range(n_sets):
        b.gen_commands_into(stage.data_ptr() + k * cmd_bytes, 1 + k, 0)
    b.gen_commands(0, 0)
    b.step(SETTLE_STEPS)
    b.sync()
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)

    def barrier():
        b.sync()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    def allgather_stats():
        _, tot = sim.allgather_stats(b.reduce_stats(), device=dev)
        return tot

    # ---- device-resident arm ------------------------------------------------------------------
    for k in range(WU):
        b.set_commands_device(stage.data_ptr() + k * cmd_bytes)
        b.step(WORLD_STEPS_PER_STEP)
        flush.zero_()
    allgather_stats()  # warms the NCCL communicator up
    torch.cuda.synchronize()
    def device_arm(periods_per_launch, sample_clocks):
        """K bench steps, `periods_per_launch` of them per step_worlds launch (the staged command blocks are
        consecutive in HBM); CUDA events bracket every launch on the batch's stream, L2 flushed in between."""
        b.enable_kernel_timing(True)
        b.kernel_time(reset=True)
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        t0 = time.perf_counter()
        k = 0
        while k < K:
            p = min(periods_per_launch, K - k)
            b.step_sequence_device(stage.data_ptr() + (WU + k) * cmd_bytes, p, WORLD_STEPS_PER_STEP)
            b.sync()
            flush.zero_()                   # L2 flush, outside the event pair
            torch.cuda.synchronize()
            k += p
        stats = allgather_stats()
        barrier()
        wall = time.perf_counter() - t0
        clk = sampler.finish() if sampler else None
        ms, n = b.kernel_time(reset=True)
        b.enable_kernel_timing(False)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), n, clk, stats

    PPL = max(1, args.periods_per_launch)
    snap = b.snapshot()
    kern_ms_max, wall_ms_max, launches, clocks, total_stats = device_arm(PPL, True)
    single = None
    if PPL > 1:     # the same K steps again, one launch per command period (what a closed-loop caller does)
        b.restore(snap)
        ms1, _, n1, _, stats1 = device_arm(1, False)
        assert stats1 == total_stats, "fused and per-period launches must produce identical worlds"
        single = {"value": W * world_size * WORLD_STEPS_PER_STEP * K / (ms1 * 1e-3), "ms_per_step": ms1 / K,
                  "launches": n1}
    b.free_snapshot(snap)
    units = W * world_size * WORLD_STEPS_PER_STEP * K
    value = units / (kern_ms_max * 1e-3)

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        n_host = min(n_sets, 16)
        host_cmd = torch.empty(n_host * cmd_bytes, dtype=torch.uint8).pin_memory()
        host_cmd.copy_(stage[: n_host * cmd_bytes].cpu())
        host_aux = torch.empty(W * 8, dtype=torch.int32).pin_memory()
        hp, ap_ = host_cmd.data_ptr(), host_aux.data_ptr()
        L, h = sim.lib(), b._h

        def whole_step(k):
            L.world_batch_set_commands(h, C.c_void_p(hp + (k % n_host) * cmd_bytes))   # H2D, pinned
            L.world_batch_step(h, WORLD_STEPS_PER_STEP)
            rc = L.world_batch_read_aux(h, C.c_void_p(ap_))                            # D2H + sync
            if rc < 0:
                raise RuntimeError("e2e step failed: %d" % rc)

        def timed(step_fn):
            for k in range(WU):
                step_fn(k)
            barrier()
            t0 = time.perf_counter()
            for k in range(K):
                step_fn(WU + k)
            barrier()                      # every chunk's last results are in host memory
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if distributed:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te[0])

        snap = b.snapshot()
        whole_s = timed(whole_step)
        e2e_s, api = whole_s, "world_batch_set_commands(host) + world_batch_step(30) + world_batch_read_aux(host)"
        nch = max(1, min(args.e2e_chunks, W))
        if nch > 1:
            stats_whole = b.reduce_stats()
            b.restore(snap)
            b.set_chunks(nch)
            ranges = [b.chunk_range(c) for c in range(nch)]

            def chunk_step(k):
                base = hp + (k % n_host) * cmd_bytes
                for c, (first, count) in enumerate(ranges):
                    # closed loop: this chunk's results of the previous step are in host memory before its next
                    # commands are submitted; the other chunks' queued work overlaps the wait
                    if L.world_batch_chunk_wait(h, c) < 0 or L.world_batch_chunk_step(
                            h, c, C.c_void_p(base + first * R * 32), WORLD_STEPS_PER_STEP,
                            C.c_void_p(ap_ + first * 32)) < 0:
                        raise RuntimeError("e2e chunk step failed")
            e2e_s = timed(chunk_step)
            assert b.reduce_stats() == stats_whole, "chunked and whole-batch stepping must produce identical worlds"
            b.set_chunks(0)
            api = ("%d chunks x [world_batch_chunk_wait + world_batch_chunk_step(host commands, 30, host aux)], "
                   "one CUDA stream per chunk" % nch)
        b.free_snapshot(snap)
        e2e = {"value": units / e2e_s, "unit": "world-steps/s", "h2d_bytes_per_step": cmd_bytes,
               "d2h_bytes_per_step": W * 32, "ms_per_step": e2e_s * 1e3 / K, "chunks": nch, "api": api,
               "whole_batch": {"value": units / whole_s, "ms_per_step": whole_s * 1e3 / K,
                               "api": "world_batch_set_commands(host) + world_batch_step(30) + "
                                      "world_batch_read_aux(host), one barrier per step"}}

    # ---- extra: the per-GPU shard of configs[4] (1,048,576 worlds over 8 GPUs = 131,072 per GPU) ---------------
    large = None
    if args.large_worlds > 0:
        WL, KL = args.large_worlds, 8
        first_l, _ = sim.shard_range(WL * world_size, rank, world_size)
        bl = sim.WorldBatch(WL, R, device=local_rank, seed=SEED, first_world_id=first_l)
        bl.gen_commands(0, 0)
        bl.step(SETTLE_STEPS)
        for k in range(3):
            bl.gen_commands(1 + k, 0)
            bl.step(WORLD_STEPS_PER_STEP)
        stage_l = torch.empty(KL * WL * R * 32, dtype=torch.uint8, device=dev)
        for k in range(KL):
            bl.gen_commands_into(stage_l.data_ptr() + k * WL * R * 32, 4 + k, 0)
        bl.sync()
        bl.enable_kernel_timing(True)
        bl.kernel_time(reset=True)
        barrier()
        bl.step_sequence_device(stage_l.data_ptr(), KL, WORLD_STEPS_PER_STEP)   # the KL periods in one launch
        ms_l, n_l = bl.kernel_time(reset=True)
        tl = torch.tensor([ms_l], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        v_l = WL * world_size * WORLD_STEPS_PER_STEP * KL / (float(tl[0]) * 1e-3)
        large = {"workload": "configs[4] shard: %d 6v6 worlds per GPU, same command stream, %d command periods "
                             "in one timed launch (working set %.0f MiB > L2 at 131,072)"
                             % (WL, KL, WL * (bpws - 16 + 32) / 2 ** 20),
                 "value": v_l, "unit": "world-steps/s", "ms_per_step": float(tl[0]) / KL,
                 "roofline_frac": (bpws * WL * WORLD_STEPS_PER_STEP) / (float(tl[0]) / KL * 1e-3) / 1e9 / measured_peak()[0]}
        bl.close()

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, el, reps = cpu_oracle_rate(R, min(W, args.ref_sample_worlds), args.cpu_seconds, threads)
        cpu = {"value": v, "unit": "world-steps/s", "cores": threads, "kind": "port",
               "sample": "%d worlds x %d world-steps x %d command batches in %.1f s; restated-Bullet CPU oracle "
                         "(oracle/ssl_oracle.c, -O2, pthreads); Bullet itself is not buildable here (SURVEY 8c)"
                         % (min(W, args.ref_sample_worlds), WORLD_STEPS_PER_STEP, reps, el)}

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_bytes = bpws * W * WORLD_STEPS_PER_STEP * K / launches
        per_launch_s = kern_ms_max * 1e-3 / launches
        achieved = per_launch_bytes / per_launch_s / 1e9
        traffic, issue = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch")
                # secondary ceiling: the kernel is issue/latency-bound, so also report warp-instructions issued per
                # second against 148 SMs x 4 schedulers x SM clock (instruction count per world-step from the
                # committed ncu capture of the 4096-world launch)
                ipw = tj["warp_instructions_per_launch"] / tj.get("world_steps_per_launch", 4096 * WORLD_STEPS_PER_STEP)
                if traffic is not None:   # scale the captured launch to this run's launch shape
                    traffic = traffic * (W * WORLD_STEPS_PER_STEP * K / launches) / tj.get("world_steps_per_launch", 1)
                peak_issue = 148 * 4 * (clocks.get("sm_max_mhz") or 1965) * 1e6
                issue = {"warp_instructions_per_world_step": ipw, "achieved_per_s": ipw * value / world_size,
                         "peak_per_s": peak_issue, "frac": ipw * value / world_size / peak_issue,
                         "source": "profiles/traffic.json (ncu smsp__inst_executed.sum)"}
            except Exception:
                traffic = None
        line = {
            "metric": "world-steps/sec (6v6, 60 Hz)", "value": value, "unit": "world-steps/s",
            "n_gpus": world_size, "steps": K, "warmup": WU, "ms_per_step": kern_ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, world_size),
            "wall_ms_per_step_incl_l2_flush": wall_ms_max / K,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches + 1,
            "periods_per_launch": PPL, "one_launch_per_period": single,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "step_worlds", "algorithmic_bytes_per_launch": per_launch_bytes,
                         "bytes_per_world_step": bpws, "kernel_ms_per_launch": kern_ms_max / launches, "issue_ceiling": issue,
                         "note": "latency/issue-bound fused kernel: state stays in registers for all "
                                 "world-steps of a launch (DRAM traffic = state once + one command block per "
                                 "period); issue_ceiling is the informative bound; see DESIGN.md"},
            "cpu_baseline": cpu,
            "large_batch": large,
            "episode_stats": total_stats,
        }
        print(json.dumps(line), flush=True)
    b.close()
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
