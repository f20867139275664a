"""Small fixed workload for ncu captures (GPU box): one chunk of the batched update at the
benchmark shape.  usage: python tools/profile_run.py [ntiles] [nb] [k] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hcore_b200 import api  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 592
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
ctx = api.Context(0)
d = bench.make_device_batch(torch, T, nb, k, dev, seed=1)
for _ in range(reps):
    bench.restore_c(d, k)
    ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"], d["crk"],
                      rk=0, maxrk=2 * k, acc=1e-8, status=d["status"])
torch.cuda.synchronize()
print("ranks", d["crk"][:4].tolist(), "status", int(d["status"].abs().sum().item()))
ctx.close()
