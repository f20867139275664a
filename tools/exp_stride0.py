"""Experiment (GPU box): stage times with every tile aliasing tile 0's factor arrays (L2-resident inputs)
versus the normal strided batch.  Separates DRAM/TLB effects from on-chip ones.  Results of the aliased
run are garbage (all tiles write the same C); only the timing is meaningful."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hcore_b200 import api
from hcore_b200.api import _ptr
T, nb, k = 2368, 1024, 64
dev = torch.device("cuda:0")
ctx = api.Context(0)
ctx.set_stage_timing(True)
d = bench.make_device_batch(torch, T, nb, k, dev, seed=1)
L = ctx._L
def run(alias):
    sA = 0 if alias else k * nb
    sC = 0 if alias else 2 * k * nb
    for it in range(3):
        bench.restore_c(d, k)
        torch.cuda.synchronize()
        rc = L.HCORE_dgemm_batched_strided(ctx._h, T, nb, -1.0, _ptr(d["au"]), _ptr(d["av"]), _ptr(d["ark"]), nb, sA,
                                           _ptr(d["bu"]), _ptr(d["bv"]), _ptr(d["brk"]), nb, sA, 1.0, _ptr(d["cu"]), _ptr(d["cv"]),
                                           _ptr(d["crk"]), nb, sC, 0, 2 * k, 1e-8, _ptr(d["status"]), None)
        torch.cuda.synchronize()
        print("alias" if alias else "normal", rc, [round(x, 2) for x in ctx.stage_ms()], flush=True)
run(False)
run(True)
