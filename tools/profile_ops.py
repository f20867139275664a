"""Small fixed workload of the secondary kernels for ncu captures (GPU box): dgemmcd, dsyrk, dgemmbdcd, d/zuncompress,
dtrsm, dpotrf at nb = 1024.  usage: python tools/profile_ops.py [ntiles]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hcore_b200 import api  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 296
nb, k = 1024, 32
dev = torch.device("cuda:0")
ctx = api.Context(0)
d = bench.make_device_batch(torch, T, nb, k, dev, seed=5)
cd = torch.rand((T, nb, nb), dtype=torch.float64, device=dev)
bd = torch.rand((T, nb, nb), dtype=torch.float64, device=dev)
for _ in range(2):
    ctx.dgemmcd_batched(d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], cd)
    ctx.dsyrk_batched(-1.0, d["au"], d["av"], d["ark"], cd)
    ctx.dgemmbdcd_batched(-1.0, d["au"], d["av"], d["ark"], bd, 1.0, cd)
    ctx.uncompress_batched(d["au"], d["ark"], d["bv"], d["brk"], cd)
    Lt = (torch.tril(cd[0].T) / 64 + torch.eye(nb, dtype=torch.float64, device=dev)).T.contiguous()[None]
    bv = d["bv"].clone()
    ctx.dtrsm_batched(141, 122, 111, 131, 1.0, Lt, bv)
    S = (cd[:8] @ cd[:8].transpose(1, 2)) / nb + torch.eye(nb, dtype=torch.float64, device=dev)
    info = torch.zeros(8, dtype=torch.int32, device=dev)
    ctx.dpotrf_batched(122, S, info)
torch.cuda.synchronize()
print("info", info.tolist())
ctx.close()
