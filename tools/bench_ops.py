"""Secondary measurements (GPU box): the other BASELINE.json configs, device-resident inputs, CUDA events.
  config 2/5: batched HCORE_dgemm sweep nb x rank      config 3: HCORE_dsyrk + HCORE_dgemmcd (dense C), nb=1024 rank 32
  config 4  : HCORE_sgemm / HCORE_zgemm, nb=1024 rank 64
usage: python tools/bench_ops.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hcore_b200 import api  # noqa: E402

dev = torch.device("cuda:0")
ctx = api.Context(0)
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, restore, reps=3, warm=2):
    for _ in range(warm):
        restore(); fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


out = {"hbm_gbs_measured": HBM}
ONLY = os.environ.get("OPS_ONLY", "")        # "", "sweep", "sz", "dense"
# ---- dgemm sweep -------------------------------------------------------------------------------------
sweep = []
for nb, k, T in [] if ONLY not in ("", "sweep") else [(512, 16, 8192), (512, 32, 8192), (512, 64, 4096), (1024, 16, 8192), (1024, 32, 8192), (1024, 64, 4736),
                 (2048, 32, 2368), (2048, 64, 2368), (1024, 128, 1184), (2048, 128, 592)]:
    d = bench.make_device_batch(torch, T, nb, k, dev, seed=3)
    fn = lambda: ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"],
                                   d["crk"], rk=0, maxrk=2 * k, acc=1e-8, status=d["status"])
    ms = timed(fn, lambda: bench.restore_c(d, k))
    r = float(d["crk"].float().mean().item())
    fl = api.gemm_flops(nb, k, k, k, int(round(r))) + 2 * (2 * k) ** 3
    sweep.append({"nb": nb, "rank": k, "tiles": T, "tiles_per_s": T / ms * 1e3, "mean_new_rank": r,
                  "gflops_effective": T / ms * 1e3 * fl / 1e9, "failed": int((d["status"] != 0).sum().item())})
    print(sweep[-1], flush=True)
    del d
    torch.cuda.empty_cache()
out["dgemm_sweep"] = sweep
# ---- sgemm / zgemm --------------------------------------------------------------------------------------
nb, k = 1024, 64
for name, dt, T, acc in () if ONLY not in ("", "sz") else (("sgemm", torch.float32, 4736, 1e-4), ("zgemm", torch.complex128, 592, 1e-8)):
    d = bench.make_device_batch(torch, T, nb, k, dev, seed=4)
    conv = {}
    for nm in ("au", "av", "bu", "bv", "cu0", "cv0", "cu", "cv"):
        x = d[nm]
        if dt == torch.complex128:
            conv[nm] = torch.complex(x, torch.roll(x, 1, dims=2) * 0.5).contiguous()
        else:
            conv[nm] = x.to(dt).contiguous()
    del d["au"], d["av"], d["bu"], d["bv"], d["cu0"], d["cv0"], d["cu"], d["cv"]

    def restore():
        conv["cu"][:, :k].copy_(conv["cu0"]); conv["cv"][:, :k].copy_(conv["cv0"]); d["crk"].fill_(k)
    call = ctx.sgemm_batched if name == "sgemm" else ctx.zgemm_batched
    fn = lambda: call(-1.0, conv["au"], conv["av"], d["ark"], conv["bu"], conv["bv"], d["brk"], 1.0, conv["cu"], conv["cv"],
                      d["crk"], rk=0, maxrk=2 * k, acc=acc, status=d["status"])
    ms = timed(fn, restore, reps=2, warm=1)
    out[name] = {"nb": nb, "rank": k, "tiles": T, "acc": acc, "tiles_per_s": T / ms * 1e3,
                 "mean_new_rank": float(d["crk"].float().mean().item()), "failed": int((d["status"] != 0).sum().item())}
    print(name, out[name], flush=True)
    del conv, d
    torch.cuda.empty_cache()
# ---- the realistic regime: tiles of one squared-exponential covariance matrix -------------------
if ONLY in ("", "cov"):
    from hcore_b200.tilegen import make_covariance_batch
    nb, maxrk, acc, reps = 1024, 128, 1e-8, 296
    b = make_covariance_batch(8, nb, maxrk, acc=acc)          # 8 distinct problems (CPU SVDs), replicated
    rep = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev).repeat((reps,) + (1,) * (x.ndim - 1)).contiguous()
    t = {n: rep(getattr(b, n)) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    cu0, cv0, crk0 = t["cu"].clone(), t["cv"].clone(), t["crk"].clone()
    T = t["au"].shape[0]
    status = torch.zeros(T, dtype=torch.int32, device=dev)

    def restore():
        t["cu"].copy_(cu0); t["cv"].copy_(cv0); t["crk"].copy_(crk0)
    fn = lambda: ctx.dgemm_batched(-1.0, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, t["cu"], t["cv"],
                                   t["crk"], rk=0, maxrk=maxrk, acc=acc, status=status)
    ms = timed(fn, restore)
    out["covariance"] = {"nb": nb, "tiles": T, "acc": acc, "tiles_per_s": T / ms * 1e3,
                         "ranks_in": [b.ark.tolist(), b.brk.tolist(), b.crk.tolist()],
                         "mean_new_rank": float(t["crk"].float().mean().item()), "failed": int((status != 0).sum().item())}
    print("covariance", out["covariance"], flush=True)
    del t, cu0, cv0
    torch.cuda.empty_cache()
# ---- the remaining tile kernels of one TLR-Cholesky step: device-resident, CUDA events --------------
if ONLY in ("", "chol"):
    nb, k = 1024, 64
    FP64 = 37.1e12
    rng = torch.Generator(device=dev); rng.manual_seed(7)
    # dtrsm: ONE lower triangular diagonal tile, T V-factors (the trsm of a Cholesky panel step)
    T = 2048
    Lt = torch.tril(torch.rand((nb, nb), dtype=torch.float64, device=dev, generator=rng) - 0.5) / 32 + torch.eye(nb, dtype=torch.float64, device=dev)
    A1 = Lt.T.contiguous()[None]                         # column-major storage of L
    BV0 = torch.rand((T, k, nb), dtype=torch.float64, device=dev, generator=rng) - 0.5
    BV = BV0.clone()
    ms = timed(lambda: ctx.dtrsm_batched(141, 122, 111, 131, 1.0, A1, BV), lambda: BV.copy_(BV0))
    byts, fl = (nb * nb // 2 + 2 * nb * k) * 8, nb * nb * k
    out["dtrsm_shared_L"] = {"nb": nb, "rank": k, "tiles": T, "calls_per_s": T / ms * 1e3, "tflops": T * fl / ms / 1e9,
                             "frac_fp64_peak": T * fl / ms / 1e9 / (FP64 / 1e12), "note": "L (4 MB) stays in L2; B in/out 1 MB per tile"}
    print("dtrsm", out["dtrsm_shared_L"], flush=True)
    # dpotrf: a single tile (the diagonal tile of a step) and a batch
    for T in (1, 64):
        X = torch.rand((T, nb, nb), dtype=torch.float64, device=dev, generator=rng) - 0.5
        S0 = (X @ X.transpose(1, 2)) / nb + torch.eye(nb, dtype=torch.float64, device=dev)
        S = S0.clone()
        info = torch.zeros(T, dtype=torch.int32, device=dev)
        ms = timed(lambda: ctx.dpotrf_batched(122, S, info), lambda: S.copy_(S0), reps=5)
        out[f"dpotrf_T{T}"] = {"n": nb, "tiles": T, "ms_per_call": ms, "calls_per_s": T / ms * 1e3,
                               "tflops": T * (nb ** 3 / 3) / ms / 1e9, "info_nonzero": int((info != 0).sum().item())}
        print("dpotrf", out[f"dpotrf_T{T}"], flush=True)
        del X, S0, S
    # tile compression (dense 1024 x 1024 -> rank <= 64): covariance-like tiles = low-rank + decaying tail
    T = 296
    gu = torch.linalg.qr(torch.rand((T, nb, 96), dtype=torch.float64, device=dev, generator=rng) - 0.5)[0]
    gv = torch.linalg.qr(torch.rand((T, nb, 96), dtype=torch.float64, device=dev, generator=rng) - 0.5)[0]
    sg = 10.0 ** (-torch.arange(96, dtype=torch.float64, device=dev) / 5.0)
    Ad = ((gu * sg) @ gv.transpose(1, 2)).transpose(1, 2).contiguous()
    Uo = torch.zeros((T, 64, nb), dtype=torch.float64, device=dev); Vo = torch.zeros_like(Uo)
    ro = torch.zeros(T, dtype=torch.int32, device=dev)
    for q in (1, 0):
        ms = timed(lambda: ctx.dcompress_batched(Ad, 1e-8, 64, Uo, Vo, ro, power_iters=q), lambda: None, reps=2, warm=1)
        out[f"dcompress_q{q}"] = {"nb": nb, "maxrk": 64, "tiles": T, "power_iters": q, "tiles_per_s": T / ms * 1e3,
                                  "mean_rank": float(ro.float().mean().item())}
        print("dcompress", out[f"dcompress_q{q}"], flush=True)
    del gu, gv, Ad, Uo, Vo
    # dgemmbdcd and uncompress: HBM-bound dense output
    T = 1024
    d = bench.make_device_batch(torch, T, nb, k, dev, seed=6)
    BD = torch.rand((T, nb, nb), dtype=torch.float64, device=dev, generator=rng) - 0.5
    CD = torch.rand((T, nb, nb), dtype=torch.float64, device=dev, generator=rng)
    ms = timed(lambda: ctx.dgemmbdcd_batched(-1.0, d["au"], d["av"], d["ark"], BD, 1.0, CD), lambda: None)
    byts = (3 * nb * nb + 2 * nb * k) * 8
    out["dgemmbdcd"] = {"nb": nb, "rank": k, "tiles": T, "calls_per_s": T / ms * 1e3, "algorithmic_GBps": T * byts / ms / 1e6,
                        "frac_of_measured_hbm": T * byts / ms / 1e6 / HBM}
    print("dgemmbdcd", out["dgemmbdcd"], flush=True)
    ms = timed(lambda: ctx.uncompress_batched(d["au"], d["ark"], d["bv"], d["brk"], CD), lambda: None)
    byts = (nb * nb + 2 * nb * k) * 8
    out["duncompress"] = {"nb": nb, "rank": k, "tiles": T, "calls_per_s": T / ms * 1e3, "algorithmic_GBps": T * byts / ms / 1e6,
                          "frac_of_measured_hbm": T * byts / ms / 1e6 / HBM}
    print("duncompress", out["duncompress"], flush=True)
    del BD, CD
    T = 512
    zu = torch.complex(d["au"][:T], d["av"][:T]).contiguous(); zv = torch.complex(d["bu"][:T], d["bv"][:T]).contiguous()
    ZD = torch.zeros((T, nb, nb), dtype=torch.complex128, device=dev)
    ms = timed(lambda: ctx.uncompress_batched(zu, d["ark"][:T], zv, d["brk"][:T], ZD), lambda: None)
    byts = (nb * nb + 2 * nb * k) * 16
    out["zuncompress"] = {"nb": nb, "rank": k, "tiles": T, "calls_per_s": T / ms * 1e3, "algorithmic_GBps": T * byts / ms / 1e6,
                          "frac_of_measured_hbm": T * byts / ms / 1e6 / HBM, "tflops": T * 8.0 * nb * nb * k / ms / 1e9}
    print("zuncompress", out["zuncompress"], flush=True)
    del d, zu, zv, ZD
    torch.cuda.empty_cache()
# ---- dense-C updates (config 3) ---------------------------------------------------------------------------
nb, k, T = 1024, 32, (2048 if ONLY in ("", "dense") else 0)
if T == 0:
    ctx.close()
    sys.exit(0)
d = bench.make_device_batch(torch, T, nb, k, dev, seed=5)
cd = torch.rand((T, nb, nb), dtype=torch.float64, device=dev)
for name, fn in (("dgemmcd", lambda: ctx.dgemmcd_batched(d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], cd)),
                 ("dsyrk", lambda: ctx.dsyrk_batched(-1.0, d["au"], d["av"], d["ark"], cd))):
    ms = timed(fn, lambda: None)
    # dsyrk reads only the blocks that touch the lower triangle and writes the whole tile (the strict upper blocks are
    # zero-filled): with 128 x 64 blocks at nb = 1024 that is 72 of 128 blocks read, 128 written
    byts = (2 * nb * nb + 4 * nb * k) * 8 if name == "dgemmcd" else (int(nb * nb * (1 + 72 / 128)) + 2 * nb * k) * 8
    out[name] = {"nb": nb, "rank": k, "tiles": T, "calls_per_s": T / ms * 1e3, "algorithmic_GBps": T * byts / ms / 1e6,
                 "frac_of_measured_hbm": T * byts / ms / 1e6 / HBM}
    print(name, out[name], flush=True)
ctx.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
