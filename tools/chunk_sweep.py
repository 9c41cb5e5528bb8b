"""e2e closed loop through host buffers: whole-batch (set_commands + step + read_aux) against the chunked
pipeline (world_batch_chunk_step / chunk_wait) for several chunk counts.  One gpurun session."""
import ctypes as C, importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sim = importlib.import_module("ssl-sim_b200")
import torch
W, R, K, WU, SPP = 4096, 12, 100, 5, 30
L = sim.lib()
cmd_bytes = W * R * 32
def fresh():
    b = sim.WorldBatch(W, R, seed=0x5351)
    b.gen_commands(0, 0); b.step(120); b.sync()
    return b
b = fresh()
stage = torch.empty((WU + K) * cmd_bytes, dtype=torch.uint8, device="cuda")
for k in range(WU + K):
    b.gen_commands_into(stage.data_ptr() + k * cmd_bytes, 1 + k, 0)
b.sync()
host_cmd = torch.empty((WU + K) * cmd_bytes, dtype=torch.uint8).pin_memory()
host_cmd.copy_(stage.cpu())
host_aux = torch.empty(W * 8, dtype=torch.int32).pin_memory()
hp, ap = host_cmd.data_ptr(), host_aux.data_ptr()
b.close()

def whole():
    b = fresh(); h = b._h
    def step(k):
        L.world_batch_set_commands(h, C.c_void_p(hp + k * cmd_bytes))
        L.world_batch_step(h, SPP)
        L.world_batch_read_aux(h, C.c_void_p(ap))
    for k in range(WU): step(k)
    b.sync(); t0 = time.perf_counter()
    for k in range(K): step(WU + k)
    b.sync(); el = time.perf_counter() - t0
    st = b.reduce_stats(); b.close()
    return W * SPP * K / el, st["checksum"]

def chunked(n):
    b = fresh(); h = b._h
    b.set_chunks(n)
    rng = [b.chunk_range(c) for c in range(n)]
    def submit(k, c):
        first, count = rng[c]
        L.world_batch_chunk_step(h, c, C.c_void_p(hp + k * cmd_bytes + first * R * 32), SPP, C.c_void_p(ap + first * 32))
    for k in range(WU):
        for c in range(n):
            L.world_batch_chunk_wait(h, c); submit(k, c)
    b.sync(); t0 = time.perf_counter()
    for k in range(K):
        for c in range(n):
            L.world_batch_chunk_wait(h, c)   # results of the chunk's previous period are in host memory
            submit(WU + k, c)
    b.sync(); el = time.perf_counter() - t0
    st = b.reduce_stats(); b.close()
    return W * SPP * K / el, st["checksum"]

for rep in range(2):
    v, cs = whole(); print("whole-batch      %.4e  checksum %d" % (v, cs), flush=True)
    for n in (2, 4, 8, 16, 32):
        v, cs = chunked(n); print("chunks=%-3d       %.4e  checksum %d" % (n, v, cs), flush=True)
