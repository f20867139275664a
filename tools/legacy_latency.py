"""Latency of the legacy one-tile HCORE_dgemm symbol (host pointers in, host pointers out) on the GPU box."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hcore_b200 import api
from hcore_b200.tilegen import make_lr_batch
for nb, k in ((1024, 64), (500, 25)):
    b = make_lr_batch(1, nb, k)
    au, av, bu, bv = (np.ascontiguousarray(x[0]) for x in (b.au, b.av, b.bu, b.bv))
    ts = []
    for it in range(12):
        cu = np.zeros((2 * k, nb)); cv = np.zeros((2 * k, nb)); cu[:k] = b.cu[0][:k]; cv[:k] = b.cv[0][:k]
        crk = np.array([float(k)])
        t0 = time.perf_counter()
        api.hcore_dgemm(111, 112, nb, nb, -1.0, au, av, np.array([float(k)]), nb, bu, bv, np.array([float(k)]), nb, 1.0,
                        cu, cv, crk, nb, 0, 2 * k, 1e-8, flops=api.FlopCounter())
        ts.append(time.perf_counter() - t0)
    print(f"legacy HCORE_dgemm nb={nb} k={k}: new rank {int(crk[0])}, median {1e3*np.median(ts[2:]):.2f} ms per call (pageable host buffers)", flush=True)
