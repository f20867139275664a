"""Small run of every kernel family for compute-sanitizer (GPU box):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hcore_b200 import api
from hcore_b200.tilegen import make_lr_batch, make_dense_batch
dev = torch.device("cuda:0")
ctx = api.Context(0)
for nb, k in ((200, 12), (257, 9), (256, 20), (300, 40), (260, 70)):    # stack widths 32, 32, 64, 128 and 256 (core matrix in L2)
    b = make_lr_batch(3, nb, k)
    t = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    st = torch.zeros(3, dtype=torch.int32, device=dev)
    ctx.dgemm_batched(-1.0, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, t["cu"], t["cv"], t["crk"], rk=0, maxrk=2 * k, acc=1e-8, status=st)
    # split phase + B dense
    t2 = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    nr = torch.zeros(3, dtype=torch.int32, device=dev)
    ctx.dgemm_batched_qr_svd(-1.0, t2["au"], t2["av"], t2["ark"], t2["bu"], t2["bv"], t2["brk"], 1.0, t2["cu"], t2["cv"], t2["crk"], nr, rk=0, maxrk=2 * k, acc=1e-8, status=st)
    ctx.dgemm_batched_ormqr(t2["cu"], t2["cv"], t2["crk"])
    Bd = torch.rand((3, nb, nb), dtype=torch.float64, device=dev) / nb
    t3 = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev) for n in ("au", "av", "cu", "cv", "ark", "crk")}
    ctx.dgemm_batched_b_dense(-1.0, t3["au"], t3["av"], t3["ark"], Bd, 1.0, t3["cu"], t3["cv"], t3["crk"], rk=0, maxrk=2 * k, acc=1e-8, status=st)
    # float and complex
    f = {n: t[n].float().contiguous() if t[n].dtype == torch.float64 else t[n] for n in t}
    f["cu"][:, k:] = 0; f["cv"][:, k:] = 0
    crk = torch.full((3,), k, dtype=torch.int32, device=dev)
    fcu = torch.from_numpy(b.cu).to(dev).float().contiguous(); fcv = torch.from_numpy(b.cv).to(dev).float().contiguous()
    ctx.sgemm_batched(-1.0, f["au"], f["av"], f["ark"], f["bu"], f["bv"], f["brk"], 1.0, fcu, fcv, crk, rk=0, maxrk=2 * k, acc=1e-4, status=st)
    z = {n: torch.complex(torch.from_numpy(getattr(b, n)).to(dev), 0.5 * torch.roll(torch.from_numpy(getattr(b, n)).to(dev), 1, dims=2)).contiguous()
         for n in ("au", "av", "bu", "bv", "cu", "cv")}
    crk = torch.full((3,), k, dtype=torch.int32, device=dev)
    if 2 * k <= 128:
      ctx.zgemm_batched(-1.0, z["au"], z["av"], t["ark"], z["bu"], z["bv"], t["brk"], 1.0, z["cu"], z["cv"], crk, rk=0, maxrk=2 * k, acc=1e-8, status=st)
    # dense C
    cd = torch.from_numpy(make_dense_batch(3, nb)).to(dev)
    ark = torch.full((3,), k, dtype=torch.int32, device=dev)
    ctx.dgemmcd_batched(t["au"], t["av"], ark, t["bu"], t["bv"], ark, cd)
    ctx.dsyrk_batched(-1.0, t["au"], t["av"], ark, cd)
    torch.cuda.synchronize()
    print("ok", nb, k, int(st.abs().sum().item()), flush=True)
ctx.close()
