"""Multi-GPU plumbing: worlds shard by contiguous range, one process per GPU, no per-step collective
(SURVEY 8e).  The only exchange is an all-gather of the per-shard episode statistics
(struct BatchStats, 8 x u64) -- torch.distributed over NCCL/NVLink on GPUs, gloo in the CPU tests."""

STAT_KEYS = ("goals_blue", "goals_yellow", "ball_out", "kicks", "touches", "contacts", "bad_worlds", "checksum")
_M64 = (1 << 64) - 1


def shard_range(total_worlds, rank, world_size):
    """GPU `rank` of `world_size` owns worlds [first, first + count): contiguous, sizes differ by at most 1."""
    if not (0 <= rank < world_size) or total_worlds < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total_worlds, world_size)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allgather_stats(stats, device=None, group=None):
    """All-gathers one BatchStats dict per rank; returns (per_rank list of dicts, total dict).
    Totals are sums modulo 2^64 (the checksum is an order-independent wrapping sum)."""
    import torch
    import torch.distributed as dist
    vals = [int(stats[k]) & _M64 for k in STAT_KEYS]
    mine = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(out, mine, group=group)
    else:
        out = [mine]
    per_rank = [{k: int(v) & _M64 for k, v in zip(STAT_KEYS, t.tolist())} for t in out]
    total = {k: sum(d[k] for d in per_rank) & _M64 for k in STAT_KEYS}
    return per_rank, total
