"""include/sslsim_multi.h on real GPUs: contiguous world ranges, no per-step collective, ncclAllGather of the episode
statistics.  The shards must equal the unsharded batch bit for bit (initial states and synthetic commands are keyed
by the global world id) and the all-gathered totals must equal the unsharded batch's reduction.  The 2-GPU case
needs two visible devices (gpurun --gpus 2) and is skipped otherwise; the CPU-side counterpart with the gloo backend
is tests/test_multirank_gloo.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _run(sim, orc, devices, total=192, R=12, periods=3):
    m = sim.WorldMulti(total, R, devices=devices)
    b = sim.WorldBatch(total, R)
    assert m.n_ranks == len(devices) and m.n_local == len(devices)
    for p in range(periods):
        m.gen_commands(p, 1); b.gen_commands(p, 1)
        m.step(30); b.step(30)
    m.sync()
    ids = list(range(0, total, 7)) + [total - 1]
    pos, vel, aux, n = m.read_worlds(ids)
    assert n == len(ids)
    bp, bv, ba = b.read_worlds(ids)
    np.testing.assert_array_equal(pos, bp)
    np.testing.assert_array_equal(vel, bv)
    np.testing.assert_array_equal(aux, ba)
    per, tot = m.allgather_stats()
    assert len(per) == len(devices)
    assert tot == b.reduce_stats()          # the checksum is an order-independent wrapping sum over worlds
    assert sum(p["contacts"] for p in per) == tot["contacts"] and tot["contacts"] > 0
    # and against the oracle on a few worlds of the last shard
    o = orc.Worlds(4, R, world_ids=[total - 4, total - 3, total - 2, total - 1])
    for p in range(periods):
        o.gen_commands(p, 1); o.step(30)
    gp, gv, ga, _ = m.read_worlds([total - 4, total - 3, total - 2, total - 1])
    np.testing.assert_array_equal(gp, o.pos); np.testing.assert_array_equal(gv, o.vel); np.testing.assert_array_equal(ga, o.aux)
    # host commands for all worlds: every shard takes its own rows
    cmds = np.zeros((total, R), sim.CMD_DTYPE)
    cmds["flags"] = sim.CMD_DRIVE
    cmds["wheel"][:, :, 0] = np.linspace(-1, 1, total, dtype=np.float32)[:, None]
    m.set_commands(cmds); b.set_commands(cmds)
    m.timer_start()
    m.step(20); b.step(20)
    ms, per_ms = m.timer_stop()
    assert ms > 0 and len(per_ms) == len(devices)
    _, tot2 = m.allgather_stats()
    assert tot2 == b.reduce_stats()
    m.close(); b.close()


def test_multi_one_device_equals_plain_batch(sim, orc):
    sim._share_torch_nccl()
    assert sim.lib().sslsim_multi_nccl_version() != b"unavailable"
    _run(sim, orc, [0])


def test_multi_two_devices_equal_unsharded_batch(sim, orc):
    if _n_gpus() < 2:
        pytest.skip("needs two visible GPUs (gpurun --gpus 2)")
    _run(sim, orc, [0, 1], total=193)       # uneven split: 97 + 96


def test_multi_all_devices(sim, orc):
    n = _n_gpus()
    if n < 4:
        pytest.skip("needs at least four visible GPUs")
    _run(sim, orc, list(range(n)), total=64 * n + 3)


def test_multi_rank_constructor_single_rank(sim, orc):
    """the one-process-per-GPU constructor (ncclCommInitRank) with a world of one rank"""
    uid = sim.WorldMulti.unique_id()
    assert len(uid) == 128
    m = sim.WorldMulti(64, 12, rank=0, n_ranks=1, device=0, nccl_id=uid)
    m.gen_commands(0, 0); m.step(30)
    per, tot = m.allgather_stats()
    b = sim.WorldBatch(64, 12)
    b.gen_commands(0, 0); b.step(30)
    assert tot == b.reduce_stats() and per[0] == tot
    m.close(); b.close()
