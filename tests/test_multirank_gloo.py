"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding + the statistics all-gather.
Each rank steps its shard with the oracle (this is a CPU test; the GPU path is covered by
test_gpu_parity.py::test_step_delta_and_shard_invariance) and the gathered totals must equal the
unsharded run."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
import oracle_binding as ob
sim = importlib.import_module("ssl-sim_b200")
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
TOTAL, R = 10, 12
first, count = sim.shard_range(TOTAL, rank, ws)
w = ob.Worlds(count, R, world0=first)
for p in range(3):
    w.gen_commands(p, 1); w.step(30)
st = {"goals_blue": int((w.aux[:, 4] & 0xFFFF).sum()), "goals_yellow": int((w.aux[:, 4] >> 16).sum()),
      "ball_out": int((w.aux[:, 5] & 0xFFFF).sum()), "kicks": int((w.aux[:, 5] >> 16).sum()),
      "touches": int(w.aux[:, 6].sum()), "contacts": int(w.aux[:, 7].astype(np.uint64).sum()),
      "bad_worlds": int((((w.aux[:, 0] >> 16) & 3) != 0).sum()), "checksum": (1 << 64) - 1 - rank}
per_rank, total = sim.allgather_stats(st)
if rank == 0:
    print("RESULT " + json.dumps({"first": [sim.shard_range(TOTAL, r, ws) for r in range(ws)], "total": total,
                                  "n": len(per_rank)}))
dist.destroy_process_group()
'''


def test_shard_range_partitions(sim):
    for total in (0, 1, 7, 4096, 1048576):
        for ws in (1, 2, 3, 8):
            spans = [sim.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gloo_allgather(orc, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][0][7:])
    assert res["n"] == 2 and res["first"] == [[0, 5], [5, 5]]
    w = orc.Worlds(10, 12)
    for p in range(3):
        w.gen_commands(p, 1)
        w.step(30)
    t = res["total"]
    assert t["contacts"] == int(w.aux[:, 7].astype(np.uint64).sum())
    assert t["kicks"] == int((w.aux[:, 5] >> 16).sum()) and t["touches"] == int(w.aux[:, 6].sum())
    assert t["checksum"] == ((1 << 64) - 1 + (1 << 64) - 2) % (1 << 64)     # wrapping u64 sum survives the int64 transport
