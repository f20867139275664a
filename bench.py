#!/usr/bin/env python
"""Benchmark of the TLR update path: batched HCORE_dgemm + recompression.

  python bench.py --gpus N --steps K --warmup W            # this library, on N B200s
  python bench.py --impl reference --gpus N ...             # the reference's CPU path, host cores

Metric (BASELINE.json): tiles/s of batched HCORE_dgemm with recompression at nb=1024,
rank 64, accuracy 1e-8 (configs[1]: 16k tiles per B200; weak scaling: 16k tiles per GPU, so
8 GPUs process the 128k tiles of configs[4]).  One "step" = one pass of the hot path over
the whole batch.  The update is in place (it overwrites C, like the reference), so the C
factors are restored from a pristine copy between steps; the restore is outside the CUDA
event pair that times the step.

One JSON line is printed by rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TLR dgemm+recompress tiles/s (nb=1024,k=64,acc=1e-8)"
FP64_PEAK_FALLBACK_TFLOPS = 37.1   # profiles/fp64_peak_r1.txt (DMMA m8n8k4, measured on this pool's B200)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=16384, help="tiles per GPU")
    ap.add_argument("--nb", type=int, default=1024)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--acc", type=float, default=1e-8)
    ap.add_argument("--e2e-tiles", type=int, default=4096, help="tiles per GPU per end-to-end (host buffer) step")
    ap.add_argument("--config", default="dgemm", choices=["dgemm", "sgemm", "zgemm", "dgemmcd", "dsyrk", "dtrsm", "dpotrf"],
                    help="dgemm = the headline (BASELINE.json configs[1]); the others are the secondary configs "
                         "(configs[2], configs[3]) and the Cholesky-step kernels, one JSON line each, N = 1")
    ap.add_argument("--cpu-sample", type=int, default=0, help="CPU baseline tiles per worker (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic batch on the device (same distribution as hcore_b200.tilegen / SURVEY 8(d))
# ------------------------------------------------------------------------------------------------
def device_orth(torch, gen, T, nb, k, dev, chunk=1024):
    """(T, k, nb) float64: each tile an nb x k matrix with orthonormal columns (column-major).
    Orthonormalised with two rounds of Cholesky-QR (a uniform random nb x k matrix is very well
    conditioned); input generation only, not part of anything timed."""
    out = torch.empty((T, k, nb), dtype=torch.float64, device=dev)
    for t0 in range(0, T, chunk):
        c = min(chunk, T - t0)
        x = torch.rand((c, k, nb), dtype=torch.float64, device=dev, generator=gen) - 0.5   # rows = columns of X
        for _ in range(2):
            g = x @ x.transpose(1, 2)                       # X^T X  (k x k)
            l = torch.linalg.cholesky(g)                     # G = L L^T,  R = L^T
            x = torch.linalg.solve_triangular(l, x, upper=False)   # (X R^-1)^T = L^-1 X^T
        out[t0:t0 + c] = x
    return out


def make_device_batch(torch, T, nb, k, dev, seed):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    sig = (10.0 ** (-8.0 * torch.arange(k, dtype=torch.float64, device=dev) / k))[None, :, None]
    d = {}
    d["au"] = device_orth(torch, gen, T, nb, k, dev)
    d["av"] = device_orth(torch, gen, T, nb, k, dev) * sig
    d["bu"] = device_orth(torch, gen, T, nb, k, dev)
    d["bv"] = device_orth(torch, gen, T, nb, k, dev) * sig
    d["cu0"] = device_orth(torch, gen, T, nb, k, dev)           # pristine C factors (k columns)
    d["cv0"] = device_orth(torch, gen, T, nb, k, dev) * sig
    d["cu"] = torch.zeros((T, 2 * k, nb), dtype=torch.float64, device=dev)
    d["cv"] = torch.zeros((T, 2 * k, nb), dtype=torch.float64, device=dev)
    d["ark"] = torch.full((T,), k, dtype=torch.int32, device=dev)
    d["brk"] = torch.full((T,), k, dtype=torch.int32, device=dev)
    d["crk"] = torch.full((T,), k, dtype=torch.int32, device=dev)
    d["status"] = torch.zeros((T,), dtype=torch.int32, device=dev)
    return d


def bind_to_gpu_numa_node(torch, index):
    """One process per GPU: run on (and first-touch the pinned host buffers from) the CPUs of the NUMA node the
    GPU hangs off, so that the end-to-end path's H2D/D2H traffic does not cross the socket interconnect.
    Returns the node id, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bus = pr.pci_bus_id
        if isinstance(bus, int):                                     # torch exposes domain / bus / device as ints
            bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), bus, getattr(pr, "pci_device_id", 0))
        base = f"/sys/bus/pci/devices/{bus.lower()}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def restore_c(d, k):
    d["cu"][:, :k].copy_(d["cu0"])
    d["cv"][:, :k].copy_(d["cv0"])
    d["crk"].fill_(k)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        if self.nv:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own HCORE_dgemm on the host cores, one process per core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wid, ntiles, nb, k, acc, use_ref, keep = args
    sys.path.insert(0, ROOT)
    from hcore_b200.tilegen import make_lr_batch
    b = make_lr_batch(ntiles + 1, nb, k, base_seed=777 + 1000 * wid)
    if use_ref:
        from oracle import ref_binding as rb
        rb.set_blas_threads(1)

        def one(t):
            return rb.hcore_gemm(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0, beta=1.0,
                                 rk=0, maxrk=2 * k, acc=acc)
    else:
        from oracle import tlr_oracle as orc

        def one(t):
            return orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0,
                                         beta=1.0, rk=0, maxrk=2 * k, acc=acc)
    one(ntiles)                       # warm-up tile
    t0 = time.perf_counter()
    res = [one(t) for t in range(ntiles)]
    dt = time.perf_counter() - t0
    ranks = [int(r[2]) for r in res]
    errs = None
    if keep:
        # the reference's own reconstruction error per tile (outside the timed loop): the same-input parity leg
        # of the GPU arm compares against it (north star: "ranks within +-1 on the same inputs")
        errs = []
        for t in range(ntiles):
            r = ranks[t]
            errs.append(float(np.linalg.norm(res[t][0][:r].T @ res[t][1][:r] - b.dense_c(t, -1.0, 1.0))))
    return dt, ranks, errs


def cpu_baseline(nb, k, acc, tiles_per_worker, workers=None, keep=False):
    """Times the reference CPU path (oracle/_ref when built, else the numpy/LAPACK port) with one
    single-threaded-BLAS process per host core (threads do not scale in this OpenBLAS)."""
    import multiprocessing as mp
    from oracle import ref_binding as rb
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    use_ref = rb.available()
    cores = workers or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(w, tiles_per_worker, nb, k, acc, use_ref, keep) for w in range(cores)])
    wall = max(r[0] for r in res)
    total = tiles_per_worker * cores
    ranks = [x for r in res for x in r[1]]
    out = {"value": total / wall, "unit": "tiles/s", "cores": cores, "kind": "reference" if use_ref else "port",
           "sample": f"{total} tiles ({tiles_per_worker} per process x {cores} single-thread-BLAS processes) of the "
                     f"same synthetic workload, nb={nb} k={k} acc={acc:g}; wall = slowest process",
           "mean_rank": float(np.mean(ranks)), "per_core": total / wall / cores}
    if keep:
        out["_ranks"] = [r[1] for r in res]
        out["_errs"] = [r[2] for r in res]
    return out


def _gen_worker(args):
    wid, ntiles, nb, k = args
    sys.path.insert(0, ROOT)
    from hcore_b200.tilegen import make_lr_batch
    b = make_lr_batch(ntiles, nb, k, base_seed=777 + 1000 * wid)
    return b.au, b.av, b.bu, b.bv, b.cu, b.cv


def same_input_parity(torch, ctx, dev, base, nb, k, acc):
    """Feeds the tiles the CPU reference just processed (regenerated from the same seeds) to the GPU path and
    compares tile by tile: |rank - reference rank| <= 1 and error <= max(acc ||C||, 1.01 x reference error)
    (+ acc when the rank is one lower: the dropped singular value), reference src/compute/hcore_dgemm.c:444-468."""
    import multiprocessing as mp
    cores, per = base["cores"], len(base["_ranks"][0])
    with mp.get_context("spawn").Pool(min(cores, 16)) as pool:
        gens = pool.map(_gen_worker, [(w, per, nb, k) for w in range(cores)])
    cat = lambda i: torch.from_numpy(np.concatenate([g[i] for g in gens])).to(dev)
    au, av, bu, bv, cu, cv = (cat(i) for i in range(6))
    T = au.shape[0]
    cu0, cv0 = cu[:, :k].clone(), cv[:, :k].clone()
    rk_ = torch.full((T,), k, dtype=torch.int32, device=dev)
    crk, st = rk_.clone(), torch.zeros((T,), dtype=torch.int32, device=dev)
    ctx.dgemm_batched(-1.0, au, av, rk_, bu, bv, rk_, 1.0, cu, cv, crk, rk=0, maxrk=2 * k, acc=acc, status=st)
    torch.cuda.synchronize()
    err, dn = tile_errors(torch, au, av, bu, bv, cu0, cv0, cu, cv, crk)
    r_ref = np.array([x for w in base["_ranks"] for x in w])
    e_ref = np.array([x for w in base["_errs"] for x in w])
    dr = crk.cpu().numpy().astype(np.int64) - r_ref
    bound = np.maximum(acc * dn, 1.01 * e_ref) + np.where(dr < 0, acc, 0.0)
    return {"tiles": int(T), "max_abs_rank_diff": int(np.abs(dr).max()), "n_rank_diff": int((dr != 0).sum()),
            "max_err_ratio": float((err / np.maximum(e_ref, 1e-300)).max()), "n_over_bound": int((err > bound).sum()),
            "max_rel_err": float((err / dn).max()), "failed_tiles": int((st != 0).sum().item()),
            "gate": "|dr| <= 1 and err <= max(acc*||C||, 1.01*ref_err) (+acc if rank is one lower)",
            "ok": bool(np.abs(dr).max() <= 1 and (err <= bound).all() and int((st != 0).sum().item()) == 0)}


def tile_errors(torch, au, av, bu, bv, cu0, cv0, cu, cv, crk, chunk=256):
    """||CU CV^T - (C - A B^T)||_F and ||C - A B^T||_F for EVERY tile (torch on the GPU; verification, untimed)."""
    T, kc, nb = cu.shape
    errs, norms = [], []
    col = torch.arange(kc, device=cu.device)[None, :, None]
    for t0 in range(0, T, chunk):
        sl = slice(t0, min(T, t0 + chunk))
        mask = (col < crk[sl].long()[:, None, None]).to(cu.dtype)
        approx = torch.bmm((cu[sl] * mask).transpose(1, 2), cv[sl] * mask)
        A = torch.bmm(au[sl].transpose(1, 2), av[sl])
        B = torch.bmm(bu[sl].transpose(1, 2), bv[sl])
        dense = torch.bmm(cu0[sl].transpose(1, 2), cv0[sl]) - torch.bmm(A, B.transpose(1, 2))
        errs.append(torch.linalg.norm((approx - dense).flatten(1), dim=1))
        norms.append(torch.linalg.norm(dense.flatten(1), dim=1))
        del approx, A, B, dense
    return torch.cat(errs).cpu().numpy(), torch.cat(norms).cpu().numpy()


def stage_flops(m, a, b, c, r):
    """Reference flop model split by pipeline stage (for per-stage roofline fractions)."""
    n = a + c
    geqrf = 2 * m * n * n - int(2 * n ** 3 / 3) + 2 * m * n + int(17 * n / 3)
    return {"stack": 4 * m * a * b, "tsqr": 2 * geqrf, "svd": 22 * n ** 3 + 2 * n ** 3 + r * n,
            "apply": (4 * r * m * n - 2 * r * n * n + 3 * r * n) +
                     (4 * m * r * n - 2 * r * n * n + 2 * r * n + m * n - n * n // 2 + n // 2)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per = args.cpu_sample or 12
    for _ in range(args.warmup):
        cpu_baseline(args.nb, args.rank, args.acc, 1)
    vals, last = [], None
    t_all = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_baseline(args.nb, args.rank, args.acc, per)
        vals.append(last["value"])
    wall = time.perf_counter() - t_all
    v = float(np.mean(vals))
    total = per * last["cores"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"HCORE_dgemm nb={args.nb} k={args.rank} acc={args.acc:g}, bounded sample of "
                                   f"{total} tiles per step on the host cores", "nb": args.nb, "rank": args.rank},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line), flush=True)


def run_secondary(args):
    """--config sgemm | zgemm | dgemmcd | dsyrk | dtrsm | dpotrf: the secondary configs of BASELINE.json through the same
    C ABI, one GPU, the same JSON schema (value = device-resident rate by CUDA events; e2e = host buffers where the
    library has a host form)."""
    import torch
    from hcore_b200 import api
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = api.Context(0)
    nb, cfg = args.nb, args.config
    k = args.rank if cfg in ("sgemm", "zgemm", "dtrsm") else 32
    hbm = 6463.3
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    sampler = ClockSampler(0)
    e2e = None
    restore = lambda: None
    if cfg in ("sgemm", "zgemm"):
        T = min(args.tiles, 4736 if cfg == "sgemm" else 1184)
        acc = 1e-4 if cfg == "sgemm" else args.acc
        d = make_device_batch(torch, T, nb, k, dev, seed=4)
        conv = {}
        for nm in ("au", "av", "bu", "bv", "cu0", "cv0", "cu", "cv"):
            x = d.pop(nm)
            conv[nm] = (torch.complex(x, torch.roll(x, 1, dims=2) * 0.5) if cfg == "zgemm" else x.to(torch.float32)).contiguous()
            del x

        def restore():
            conv["cu"][:, :k].copy_(conv["cu0"]); conv["cv"][:, :k].copy_(conv["cv0"]); d["crk"].fill_(k)
        call = ctx.sgemm_batched if cfg == "sgemm" else ctx.zgemm_batched
        fn = lambda: call(-1.0, conv["au"], conv["av"], d["ark"], conv["bu"], conv["bv"], d["brk"], 1.0, conv["cu"], conv["cv"],
                          d["crk"], rk=0, maxrk=2 * k, acc=acc, status=d["status"])
        unit, work = "tiles/s", f"batched HCORE_{cfg}, {T} tiles, nb={nb}, rank {k}, acc {acc:g} (BASELINE.json configs[3])"
        flops_unit = (8.0 if cfg == "zgemm" else 1.0) * (api.gemm_flops(nb, k, k, k, 93) + 2 * (2 * k) ** 3) / (4.0 if cfg == "zgemm" else 1.0)
        es = 4 if cfg == "sgemm" else 16
        bytes_unit = es * nb * (6 * k + 2 * 93)
    elif cfg in ("dgemmcd", "dsyrk"):
        T = min(args.tiles, 2048)
        d = make_device_batch(torch, T, nb, k, dev, seed=5)
        cd = torch.rand((T, nb, nb), dtype=torch.float64, device=dev)
        fn = (lambda: ctx.dgemmcd_batched(d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], cd)) if cfg == "dgemmcd" else \
             (lambda: ctx.dsyrk_batched(-1.0, d["au"], d["av"], d["ark"], cd))
        unit, work = "calls/s", f"batched HCORE_{cfg}, {T} dense {nb} x {nb} tiles, rank {k} (BASELINE.json configs[2])"
        flops_unit = 2.0 * nb * nb * k + 4.0 * nb * k * k
        bytes_unit = (2 * nb * nb + 4 * nb * k) * 8 if cfg == "dgemmcd" else (int(nb * nb * (1 + 72 / 128)) + 2 * nb * k) * 8
    elif cfg == "dtrsm":
        T = min(args.tiles, 2048)
        g = torch.Generator(device=dev); g.manual_seed(7)
        Lt = torch.tril(torch.rand((nb, nb), dtype=torch.float64, device=dev, generator=g) - 0.5) / 32 + torch.eye(nb, dtype=torch.float64, device=dev)
        A1 = Lt.T.contiguous()[None]
        B0 = torch.rand((T, k, nb), dtype=torch.float64, device=dev, generator=g) - 0.5
        B = B0.clone()
        restore = lambda: B.copy_(B0)
        fn = lambda: ctx.dtrsm_batched(141, 122, 111, 131, 1.0, A1, B)
        unit, work = "calls/s", f"batched HCORE_dtrsm (Left, Lower, NoTrans), one shared {nb} x {nb} triangle, {T} factors of rank {k}"
        flops_unit, bytes_unit = float(nb) * nb * k, 2 * nb * k * 8
    else:
        T = min(args.tiles, 64)
        X = torch.rand((T, nb, nb), dtype=torch.float64, device=dev) - 0.5
        S0 = (X @ X.transpose(1, 2)) / nb + torch.eye(nb, dtype=torch.float64, device=dev)
        S = S0.clone()
        info = torch.zeros(T, dtype=torch.int32, device=dev)
        restore = lambda: S.copy_(S0)
        fn = lambda: ctx.dpotrf_batched(122, S, info)
        unit, work = "calls/s", f"batched HCORE_dpotrf (Lower), {T} dense {nb} x {nb} tiles"
        flops_unit, bytes_unit = nb ** 3 / 3.0, nb * nb * 8
    for _ in range(max(args.warmup, 3)):
        restore(); fn()
    torch.cuda.synchronize()
    sampler.start()
    l0, ms = ctx.launches, 0.0
    for _ in range(args.steps):
        restore()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    launches = ctx.launches - l0
    clocks = sampler.stop()
    ms /= args.steps
    value = T / (ms * 1e-3)
    tf, gbs = value * flops_unit / 1e12, value * bytes_unit / 1e9
    fp_frac, hbm_frac = tf / FP64_PEAK_FALLBACK_TFLOPS, gbs / hbm
    bound = "fp64" if fp_frac >= hbm_frac else "hbm"
    roofline = {"bound": bound, "achieved": tf if bound == "fp64" else gbs, "peak": FP64_PEAK_FALLBACK_TFLOPS if bound == "fp64" else hbm,
                "unit": "TFLOP/s" if bound == "fp64" else "GB/s", "frac": max(fp_frac, hbm_frac), "traffic": None,
                "note": f"model flops {flops_unit:.4g} and algorithmic bytes {bytes_unit:.4g} per unit; FP64 peak 37.1 TF (profiles/fp64_peak_r1.txt), "
                        f"HBM {hbm:.0f} GB/s (MEASURED_PEAKS.json); the binding roof is the one with the larger fraction "
                        f"(fp64 {fp_frac:.3f}, hbm {hbm_frac:.3f})"}
    if cfg in ("sgemm", "zgemm") and not args.no_e2e:
        Te = min(1024 if cfg == "zgemm" else 2048, T)
        host = {nm: conv[nm][:Te].cpu().pin_memory() for nm in ("au", "av", "bu", "bv")}
        host["cu"] = torch.zeros((Te, 2 * k, nb), dtype=conv["cu"].dtype).pin_memory()
        host["cv"] = torch.zeros_like(host["cu"]).pin_memory()
        cu0h, cv0h = conv["cu0"][:Te].cpu(), conv["cv0"][:Te].cpu()
        ark = torch.full((Te,), k, dtype=torch.int32).pin_memory(); brk = ark.clone().pin_memory(); crk = ark.clone().pin_memory()
        sth = torch.zeros((Te,), dtype=torch.int32).pin_memory()
        times = []
        for it in range(2 + args.steps):
            host["cu"][:, :k].copy_(cu0h); host["cv"][:, :k].copy_(cv0h); crk.fill_(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.dgemm_batched_host(-1.0, host["au"], host["av"], ark, host["bu"], host["bv"], brk, 1.0, host["cu"], host["cv"], crk,
                                   rk=0, maxrk=2 * k, acc=acc, status=sth)
            if it >= 2:
                times.append(time.perf_counter() - t0)
        es = 4 if cfg == "sgemm" else 16
        e2e = {"value": Te / float(np.mean(times)), "unit": unit, "h2d_bytes_per_step": Te * 6 * k * nb * es,
               "d2h_bytes_per_step": Te * 2 * 2 * k * nb * es, "tiles_per_step": Te, "failed_tiles": int((sth != 0).sum().item()),
               "note": f"HCORE_{cfg}_batched_host, pinned host buffers, three-stream pipeline"}
    line = {"metric": f"HCORE_{cfg} batched {unit} (nb={nb}, rank {k})", "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"sgemm": "f32 in/out, f64 arithmetic", "zgemm": "c128"}.get(cfg, "f64"), "data": "synthetic",
            "config": {"workload": work, "nb": nb, "rank": k, "units_per_step": T, "l2": "inputs exceed L2"},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks, "e2e": e2e}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config != "dgemm":
        run_secondary(args)
        return

    import torch
    import torch.distributed as dist
    from hcore_b200 import api
    from hcore_b200.scheduler import TileScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hcore_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's banner out of stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)

    nb, k, acc = args.nb, args.rank, args.acc
    T_local = args.tiles
    total_tiles = T_local * world
    sched = TileScheduler(total_tiles, world, rank)
    assert sched.count == T_local
    ctx = api.Context(local)
    d = make_device_batch(torch, T_local, nb, k, dev, seed=20260 + sched.first)
    torch.cuda.synchronize()

    def hot_path():
        ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"],
                          d["crk"], rk=0, maxrk=2 * k, acc=acc, status=d["status"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        restore_c(d, k)
        hot_path()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = ctx.launches
    wall0 = time.perf_counter()
    for s in range(args.steps):
        restore_c(d, k)               # untimed: puts C back to its pristine rank-k state
        barrier()
        ev0[s].record()
        hot_path()
        ev1[s].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    ms_local = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    tms = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms = float(tms.item())
    ms_per_step = total_ms / args.steps
    value = total_tiles / (ms_per_step * 1e-3)

    # results of the last step: ranks / status gathered with NCCL (the only collective on this path)
    # through the library's own C entry point (hcore_b200_gather_status: ncclAllGather on a raw ncclComm_t); the
    # torch.distributed gather is only the fallback when NCCL's C API cannot be reached from this process
    try:
        ranks_all, stat_all = sched.gather_results_c(ctx, d["crk"], d["status"])
        gather_via = "hcore_b200_gather_status (C ABI, ncclAllGather)" if world > 1 else "hcore_b200_gather_status (single rank: copy)"
    except Exception as e:  # pragma: no cover
        ranks_all, stat_all = sched.gather_results(d["crk"], d["status"])
        gather_via = f"torch.distributed all_gather (C-ABI gather unavailable: {e!r})"
    ranks_np = ranks_all.cpu().numpy()
    nfail = int((stat_all != 0).sum().item())
    mean_rank = float(ranks_np.mean())

    # accuracy of EVERY local tile of the last timed step (dense check with torch on the GPU, untimed)
    err_all, dn_all = tile_errors(torch, d["au"], d["av"], d["bu"], d["bv"], d["cu0"], d["cv0"], d["cu"], d["cv"], d["crk"])
    acc_stats = torch.tensor([float((err_all / dn_all).max()), float((err_all / dn_all > acc).sum()),
                              float(err_all.max())], dtype=torch.float64, device=dev)
    if world > 1:
        mx = acc_stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = acc_stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        acc_stats = torch.stack([mx[0], sm[1], mx[2]])
    accuracy = {"tiles_checked": total_tiles, "max_rel_err": float(acc_stats[0].item()),
                "n_over_acc": int(acc_stats[1].item()), "max_abs_err_fro": float(acc_stats[2].item()), "acc": acc}
    errs = [float(x) for x in err_all[:4]]

    # per-stage device times (extra pass with event timing between the kernels; not part of `value`)
    restore_c(d, k)
    ctx.set_stage_timing(True)
    hot_path()
    torch.cuda.synchronize()
    stage_ms = ctx.stage_ms()
    ctx.set_stage_timing(False)
    r_model = int(round(mean_rank))
    fl = stage_flops(nb, k, k, k, r_model)
    flops_tile = api.gemm_flops(nb, k, k, k, r_model) + 2 * (2 * k) ** 3

    # FP64 peak calibration: cuBLAS DGEMM 8192^3 (calibration only) next to the DMMA micro-benchmark figure
    a64 = torch.rand((8192, 8192), dtype=torch.float64, device=dev)
    b64 = torch.rand((8192, 8192), dtype=torch.float64, device=dev)
    torch.matmul(a64, b64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        torch.matmul(a64, b64)
    e1.record()
    torch.cuda.synchronize()
    cublas_tflops = 3 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a64, b64
    peak_tflops = max(FP64_PEAK_FALLBACK_TFLOPS, cublas_tflops)

    names = ["stack", "tsqr", "svd", "apply"]
    nchunks = max(1, launches // (args.steps * 4)) if launches else 1
    stages = {}
    for nm, ms in zip(names, stage_ms):
        ach = fl[nm] * T_local / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        stages[nm] = {"ms": ms, "share": ms / max(sum(stage_ms), 1e-9), "model_tflops": ach, "frac_fp64_peak": ach / peak_tflops}
    dom = max(names, key=lambda nm: stages[nm]["ms"])
    # DRAM bytes of the dominant kernel per launch: dram__bytes_read.sum + dram__bytes_write.sum of one
    # `ncu --set full` capture (per tile, profiles/roofline_traffic.json) x the tiles one launch processes
    traffic, traffic_src, pipe_ncu = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = tj.get(f"tlr_{dom}_kernel")
        if ent and (nb, k) == (tj.get("nb"), tj.get("rank")):
            traffic = ent["dram_bytes_per_tile"] * (T_local / nchunks)
            traffic_src = tj.get("source")
            pipe_ncu = ent.get("fp64_pipe_active_pct")
            for nm in names:                     # ncu-measured FP64 pipe utilisation of each stage's kernel (committed capture)
                e2 = tj.get(f"tlr_{nm}_kernel")
                if e2 and "fp64_pipe_active_pct" in e2:
                    stages[nm]["ncu_fp64_pipe_active_pct"] = e2["fp64_pipe_active_pct"]
                    stages[nm]["ncu_dram_bytes_per_tile"] = e2["dram_bytes_per_tile"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": f"tlr_{dom}_kernel", "achieved": stages[dom]["model_tflops"],
                "peak": peak_tflops, "unit": "TFLOP/s", "frac": stages[dom]["model_tflops"] / peak_tflops,
                "traffic": traffic,
                "note": "FP64: peak = max(DMMA m8n8k4 micro-benchmark 37.1 TF (profiles/fp64_peak_r1.txt), cuBLAS DGEMM "
                        f"8192^3 measured in this run = {cublas_tflops:.1f} TF); MEASURED_PEAKS.json has no FP64 entry. "
                        "achieved = reference flop model of the stage x tiles / CUDA-event time of the stage's launches",
                "traffic_source": traffic_src, "launches_per_step": nchunks,
                "ncu_fp64_pipe_active_pct": pipe_ncu,
                "ncu_note": "sm__pipe_shared_cycles_active (the FP64 pipe DMMA and DFMA share) of the committed ncu --set full capture: what "
                            "the kernel keeps busy, counting every FP64 instruction -- `frac` counts only the reference's model flops",
                "pipeline_frac": value / world * flops_tile / 1e12 / peak_tflops}

    line = {"metric": METRIC, "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batched HCORE_dgemm, {T_local} tiles per GPU, nb={nb}, rank {k}, acc {acc:g}, "
                                   f"alpha=-1 beta=1 (BASELINE.json configs[1]; x{world} GPUs = {total_tiles} tiles)",
                       "tiles_per_gpu": T_local, "nb": nb, "rank": k, "maxrk": 2 * k, "acc": acc,
                       "parallelism": f"tile-sharded x{world}", "l2": "inputs (>60 GB per GPU) exceed L2",
                       "c_restore_between_steps": "untimed device copy of the pristine C factors"},
            "gflops_effective": value * flops_tile / 1e9, "flops_per_tile_model": flops_tile,
            "mean_new_rank": mean_rank, "failed_tiles": nfail, "status_gather": gather_via, "sample_abs_err_fro": errs, "accuracy": accuracy,
            "gpu_launches": launches, "stages": stages, "roofline": roofline, "clocks": clocks,
            "wall_s_timed_region": wall}

    # ---- end to end through the C ABI with HOST buffers (H2D + D2H inside the timed region) ----
    # every rank runs it on its own GPU at the same time; the time is the max over ranks
    if not args.no_e2e:
        Te = min(args.e2e_tiles, T_local)                   # the same tile count per GPU at every N
        host = {}
        for nm in ("au", "av", "bu", "bv"):
            host[nm] = d[nm][:Te].cpu().pin_memory()
        host["cu"] = torch.zeros((Te, 2 * k, nb), dtype=torch.float64).pin_memory()
        host["cv"] = torch.zeros((Te, 2 * k, nb), dtype=torch.float64).pin_memory()
        cu0h, cv0h = d["cu0"][:Te].cpu(), d["cv0"][:Te].cpu()
        ark = torch.full((Te,), k, dtype=torch.int32).pin_memory()
        brk = torch.full((Te,), k, dtype=torch.int32).pin_memory()
        crk = torch.full((Te,), k, dtype=torch.int32).pin_memory()
        sth = torch.zeros((Te,), dtype=torch.int32).pin_memory()
        times = []
        for it in range(args.warmup + args.steps):
            host["cu"][:, :k].copy_(cu0h)
            host["cv"][:, :k].copy_(cv0h)
            crk.fill_(k)
            barrier()
            t0 = time.perf_counter()
            ctx.dgemm_batched_host(-1.0, host["au"], host["av"], ark, host["bu"], host["bv"], brk, 1.0,
                                   host["cu"], host["cv"], crk, rk=0, maxrk=2 * k, acc=acc, status=sth)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        te = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
        nf = torch.tensor([float((sth != 0).sum().item())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(nf, op=dist.ReduceOp.SUM)
        # bytes that actually cross PCIe (the library copies only the columns that carry data: Crk columns of
        # C in, max new rank of each slice out)
        # bytes that actually cross PCIe (the library copies only the columns that carry data: Crk columns of C in,
        # min(maxrk, Ark + Crk) columns of C out -- fixed before the new ranks are known, so nothing waits for them)
        per_tile_in = (4 * k + 2 * k) * nb * 8
        out_bytes = 2 * Te * nb * min(2 * k, 2 * k) * 8
        line["e2e"] = {"value": world * Te / float(te.item()), "unit": "tiles/s",
                       "h2d_bytes_per_step": world * (Te * per_tile_in + 3 * 4 * Te),
                       "d2h_bytes_per_step": world * (out_bytes + 2 * 4 * Te),
                       "tiles_per_step": world * Te, "failed_tiles": int(nf.item()), "rank0_numa_node": numa,
                       "note": "HCORE_dgemm_batched_host on every rank at once: pinned host buffers, slices of 148 tiles "
                               "through three staging buffers on three streams (H2D / kernels / D2H, event-chained, no "
                               "host synchronisation inside the loop); host wall clock, max over ranks"}
        # the transfer roofline of the end-to-end path: bytes moved / time against the raw concurrent H2D + D2H ceiling
        # of this box measured with the same slice sizes (tools/pcie_ceiling.py -> profiles/pcie_ceiling_n*.json)
        e = line["e2e"]
        ach = (e["h2d_bytes_per_step"] + e["d2h_bytes_per_step"]) / float(te.item()) / 1e9
        ceil_gbs, src = None, None
        for cand in (os.path.join(ROOT, "profiles", f"pcie_ceiling_n{world}.json"),):
            if os.path.exists(cand):
                pj = json.load(open(cand))
                # the H2D direction is the one that binds: the ceiling in tiles/s is the H2D rate / bytes in per tile
                ceil_gbs = pj["both_h2d_gbs"] * (1.0 + e["d2h_bytes_per_step"] / e["h2d_bytes_per_step"])
                src = os.path.relpath(cand, ROOT)
        note = ("ceiling = measured H2D rate with D2H running concurrently, scaled to this call's H2D:D2H byte mix (H2D is the "
                "binding direction)")
        if world > 1:
            note += ("; the N > 1 ceilings were measured on ONE 8-GPU box (profiles/pcie_ceiling_8gpu_box_n*.json): on a box whose "
                     "host moves more per GPU the fraction can exceed 1")
        e["roofline"] = {"bound": "pcie", "achieved_gbs": ach, "ceiling_gbs": ceil_gbs, "frac": (ach / ceil_gbs) if ceil_gbs else None,
                         "ceiling_source": src, "note": note}
        del host
    if rank == 0 and world == 1:
        # ---- CPU baseline: the reference on the host cores ----
        if not args.no_cpu_baseline:
            try:
                base = cpu_baseline(nb, k, acc, args.cpu_sample or 24, keep=True)
                line["parity"] = same_input_parity(torch, ctx, dev, base, nb, k, acc)
                line["cpu_baseline"] = {kk: v for kk, v in base.items() if not kk.startswith("_")}
                pc = line["cpu_baseline"]["per_core"]
                line["host_core_equivalents"] = {"device_resident": value / pc, "end_to_end": line.get("e2e", {}).get("value", 0.0) / pc,
                                                 "note": "this B200 = that many host cores of the reference at the measured tiles/s/core"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
