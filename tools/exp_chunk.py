"""Experiment (GPU box): throughput versus workspace chunk size (wave quantisation of the per-stage launches)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hcore_b200 import api
nb, k = 1024, 64
dev = torch.device("cuda:0")
ctx = api.Context(0)
T = 8880
d = bench.make_device_batch(torch, T, nb, k, dev, seed=1)
for chunk in (2368, 2220, 2664, 1776, 4440, 1480):
    ctx.set_chunk(chunk)
    ts = []
    for it in range(3):
        bench.restore_c(d, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"], d["crk"],
                          rk=0, maxrk=2 * k, acc=1e-8, status=d["status"])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("chunk", chunk, "tiles/s", round(T / min(ts) * 1e3), [round(x, 1) for x in ts], flush=True)
