"""Per-phase cycle shares inside the QR and SVD kernels (GPU box), from the instrumented build
(`make -C hcore_b200/csrc prof`).  One designated thread per CTA accumulates clock64() deltas
between checkpoints, so the numbers are CTA-resident cycles per phase, summed over CTAs.
usage: HCORE_B200_LIB=hcore_b200/lib/libhcore_b200_prof.so python tools/phase_prof.py [ntiles] [nb] [k]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HCORE_B200_LIB", os.path.join(ROOT, "hcore_b200", "lib", "libhcore_b200_prof.so"))
import bench  # noqa: E402
from hcore_b200 import api, _lib  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
ctx = api.Context(0)
L = _lib.load()
L.hcore_b200_prof_read.argtypes = [C.c_void_p, C.c_int, C.c_int]
d = bench.make_device_batch(torch, T, nb, k, dev, seed=1)
buf = (C.c_ulonglong * 64)()


def run():
    bench.restore_c(d, k)
    ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"], d["crk"],
                      rk=0, maxrk=2 * k, acc=1e-8, status=d["status"])
    torch.cuda.synchronize()


run()
L.hcore_b200_prof_read(buf, 64, 1)
run()
L.hcore_b200_prof_read(buf, 64, 1)
v = list(buf)
TSQR = {0: "load columns", 1: "chain: records of earlier groups (TMA wait + DMMA)", 2: "fence + sync after chain",
        3: "chain: records of the group's earlier warps (wait + DMMA)", 4: "warp_factor (own 128 x 8 sub-panel)",
        5: "publish + wait until the slot is consumed"}
PF = {}
SVD = {10: "core product T'", 11: "column norms", 12: "QRCP", 13: "transpose", 14: "Jacobi", 15: "sigma/sort/truncate",
       16: "U_r scatter", 17: "W products"}
for name, tab, cnt in (("tlr_tsqr_kernel", TSQR, 9), ("  inside panel_factor", PF, 9), ("tlr_svd_kernel", SVD, 29)):
    tot = sum(v[s] for s in tab)
    n = max(v[cnt], 1)
    print(f"== {name}: {n} CTAs, {tot / n:,.0f} cycles per CTA in the listed phases")
    for s, label in tab.items():
        print(f"   {label:58s} {v[s] / n:12,.0f} cyc/CTA  {100.0 * v[s] / max(tot, 1):5.1f} %")
if v[53]:
    print("== jacobi_pair_step (warp 0): %d pair steps; cycles per pair step: Gram %.0f, rotations %.0f, apply %.0f; per step: "
          "pair step + overhead %.0f, barrier wait %.0f" % (v[53], v[50] / v[53], v[51] / v[53], v[52] / v[53], v[54] / v[53], v[55] / v[53]))
if v[45]:
    print("== qrcp_step (thread 0): %d steps; cycles per step: barrier wait %.0f, winner %.0f, householder + dot %.0f, update %.0f, "
          "norms + publish %.0f" % (v[45], v[40] / v[45], v[41] / v[45], v[42] / v[45], v[43] / v[45], v[44] / v[45]))
print(f"QRCP steps per tile {v[20] / max(v[29], 1):.1f}, Jacobi sweeps per tile {v[21] / max(v[29], 1):.2f}")
ctx.close()
