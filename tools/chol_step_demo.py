"""Device-resident TLR Cholesky built ONLY from this library's batched tile kernels (GPU box):

    HCORE_dcompress (every off-diagonal tile, dense -> U V^T)
    for k:  HCORE_dpotrf(A_kk)                                   dense diagonal tile
            HCORE_dtrsm (V_ik <- L_kk^-1 V_ik, i > k)            one shared triangle, ragged ranks
            HCORE_dsyrk (A_ii -= A_ik A_ik^T, i > k)             dense diagonal tiles
            HCORE_dgemm (A_ij -= A_ik A_jk^T, i > j > k)         low-rank update + recompression, pointer arrays

The tiles never leave HBM between the kernels (the host only builds pointer tables); this is the regime in which the
device-resident throughput of bench.py -- not the PCIe-bound end-to-end figure -- is what a caller sees.
The matrix is a squared-exponential covariance matrix over Morton-ordered points (HiCMA's synthetic problem), compressed
tile by tile to `acc`; the result is checked against the dense matrix: ||A - L L^T||_F / ||A||_F.

usage: python tools/chol_step_demo.py [nt nb maxrk acc] [out.json]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hcore_b200 import api  # noqa: E402
from hcore_b200.tilegen import _morton_points  # noqa: E402

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
maxrk = int(sys.argv[3]) if len(sys.argv) > 3 else 128
acc = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
outp = sys.argv[5] if len(sys.argv) > 5 else None
dev = torch.device("cuda:0")
N = nt * nb

# ---- the matrix and its TLR form (host, once) ----------------------------------------------------------------------
rng = np.random.default_rng(7)
pts = _morton_points(N, rng)
dd = torch.from_numpy(pts).to(dev)
K = torch.exp(-(torch.cdist(dd, dd) / 0.1) ** 2) + 1e-2 * torch.eye(N, dtype=torch.float64, device=dev)   # SPD, well inside acc
ctx = api.Context(0)
# the off-diagonal tiles are compressed ON THE DEVICE (HCORE_dcompress_batched_strided): tile (i, k) at index k * nt + i
idx = [(i, k) for k in range(nt) for i in range(k + 1, nt)]
Adense = torch.stack([K[i * nb:(i + 1) * nb, k * nb:(k + 1) * nb].T.contiguous() for (i, k) in idx])     # column-major tiles
Uc = torch.zeros((len(idx), maxrk, nb), dtype=torch.float64, device=dev); Vc = torch.zeros_like(Uc)
rkc = torch.zeros(len(idx), dtype=torch.int32, device=dev); stc = torch.zeros_like(rkc)
torch.cuda.synchronize(); tc0 = time.perf_counter()
ctx.dcompress_batched(Adense, acc, maxrk // 4, Uc, Vc, rkc, status=stc)
torch.cuda.synchronize(); t_compress = time.perf_counter() - tc0
assert int(stc.abs().sum().item()) == 0
Ud = torch.zeros((nt * nt, maxrk, nb), dtype=torch.float64, device=dev); Vd = torch.zeros_like(Ud)
rkd = torch.ones(nt * nt, dtype=torch.int32, device=dev)
sel = torch.tensor([k * nt + i for (i, k) in idx], device=dev)
Ud[sel], Vd[sel], rkd[sel] = Uc, Vc, rkc
del Adense, Uc, Vc
Dd = torch.stack([K[i * nb:(i + 1) * nb, i * nb:(i + 1) * nb].T.contiguous() for i in range(nt)])
print(f"N = {N} ({nt} x {nt} tiles of {nb}), acc {acc:g}: {len(idx)} tiles compressed on the device in {t_compress * 1e3:.1f} ms, "
      f"ranks min/mean/max = {int(rkc.min())}/{float(rkc.float().mean()):.1f}/{int(rkc.max())}")
info = torch.zeros(nt, dtype=torch.int32, device=dev)
L = ctx._L
import ctypes as C  # noqa: E402
vp = lambda t: C.c_void_p(t.data_ptr())
tile_bytes = maxrk * nb * 8
ubase, vbase = Ud.data_ptr(), Vd.data_ptr()
launch0 = ctx.launches
torch.cuda.synchronize()
t0 = time.perf_counter()
ngemm = 0
for k in range(nt):
    ctx.dpotrf_batched(api.HCORE_Lower, Dd[k:k + 1], info[k:k + 1])
    m = nt - k - 1
    if m == 0:
        break
    col = slice(k * nt + k + 1, k * nt + nt)                                       # tiles (i, k), i > k: contiguous
    ctx.dtrsm_batched(api.HCORE_Left, api.HCORE_Lower, api.HCORE_NoTrans, api.HCORE_NonUnit, 1.0, Dd[k:k + 1], Vd[col], rkd[col])
    ctx.dsyrk_batched(-1.0, Ud[col], Vd[col], rkd[col], Dd[k + 1:])
    # A_ij -= A_ik A_jk^T for i > j > k: one pointer-array call over all pairs
    ii, jj = np.tril_indices(m, -1)
    if len(ii):
        ti = torch.from_numpy((k * nt + k + 1 + ii).astype(np.int64)).to(dev)      # A = tile (i, k)
        tj = torch.from_numpy((k * nt + k + 1 + jj).astype(np.int64)).to(dev)      # B = tile (j, k)
        tc = torch.from_numpy(((k + 1 + jj) * nt + k + 1 + ii).astype(np.int64)).to(dev)   # C = tile (i, j)
        pAU, pAV = ubase + ti * tile_bytes, vbase + ti * tile_bytes
        pBU, pBV = ubase + tj * tile_bytes, vbase + tj * tile_bytes
        pCU, pCV = ubase + tc * tile_bytes, vbase + tc * tile_bytes
        ark, brk, crk = rkd[ti].contiguous(), rkd[tj].contiguous(), rkd[tc].contiguous()
        st = torch.zeros(len(ii), dtype=torch.int32, device=dev)
        rc = L.HCORE_dgemm_batched(ctx._h, len(ii), nb, -1.0, vp(pAU), vp(pAV), vp(ark), nb, vp(pBU), vp(pBV), vp(brk), nb, 1.0,
                                   vp(pCU), vp(pCV), vp(crk), nb, 0, maxrk, acc, vp(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, ctx._L.hcore_b200_last_error(ctx._h)
        rkd[tc] = crk
        assert int(st.abs().sum().item()) == 0, st[st != 0][:8]
        ngemm += len(ii)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
assert int(info.abs().sum().item()) == 0, info

# ---- check: assemble L (dense) and compare L L^T with the matrix -----------------------------------------------------
Lm = torch.zeros((N, N), dtype=torch.float64, device=dev)
for k in range(nt):
    Lm[k * nb:(k + 1) * nb, k * nb:(k + 1) * nb] = torch.tril(Dd[k].T)
    for i in range(k + 1, nt):
        r = int(rkd[k * nt + i].item())
        Lm[i * nb:(i + 1) * nb, k * nb:(k + 1) * nb] = Ud[k * nt + i, :r].T @ Vd[k * nt + i, :r]
err = (torch.linalg.norm(Lm @ Lm.T - K) / torch.linalg.norm(K)).item()
res = {"N": N, "nt": nt, "nb": nb, "maxrk": maxrk, "acc": acc, "seconds": dt, "compress_seconds": t_compress, "tiles_compressed": len(idx), "gemm_tile_updates": ngemm,
       "kernel_launches": ctx.launches - launch0, "rel_err_LLt": err, "final_rank_mean": float(rkd[rkd > 1].float().mean().item()),
       "final_rank_max": int(rkd.max().item()),
       "note": "compress / potrf / trsm / syrk / gemm all through libhcore_b200 on device-resident tiles; wall clock incl. host-side pointer tables"}
print(json.dumps(res))
if outp:
    json.dump(res, open(outp, "w"), indent=1)
assert err < 100 * acc, err
ctx.close()
