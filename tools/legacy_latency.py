"""Throughput of the LEGACY single-tile drop-in symbol (HCORE_dgemm with host pointers, one tile per call) when it
is called the way a task runtime calls the reference: from W worker threads at once, one tile per call
(HiCMA over StarPU / PaRSEC: one worker per core).  Every calling thread owns its context and stream, so the calls
overlap on the GPU.  Next to it: the compiled reference (oracle/_ref) called the same way from W processes
(threads do not scale in this OpenBLAS build).  usage: python tools/legacy_latency.py [nb k tiles_per_worker]"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hcore_b200 import api  # noqa: E402
from hcore_b200.tilegen import make_lr_batch  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
per = int(sys.argv[3]) if len(sys.argv) > 3 else 24
b = make_lr_batch(8, nb, k, base_seed=3)


def worker(n, out, wid):
    fl = api.FlopCounter()
    ranks = []
    for i in range(n):
        t = (wid + i) % 8
        CU = np.zeros((2 * k, nb)); CV = np.zeros((2 * k, nb)); CU[:k] = b.cu[t][:k]; CV[:k] = b.cv[t][:k]
        Crk = np.array([float(k)])
        api.hcore_dgemm(111, 112, nb, nb, -1.0, b.au[t], b.av[t], np.array([float(k)]), nb, b.bu[t], b.bv[t], np.array([float(k)]), nb,
                        1.0, CU, CV, Crk, nb, 0, 2 * k, 1e-8, None, fl)
        ranks.append(int(Crk[0]))
    out[wid] = ranks


def persistent_worker(wid, start, stop, counts):
    """A runtime worker: lives for the whole run (its context is created once), processes tiles until told to stop."""
    out = {}
    worker(2, out, wid)                                 # first calls of the thread: context + workspace allocation
    start.wait()
    n = 0
    while not stop.is_set():
        worker(1, out, wid + n)
        n += 1
    counts[wid] = n


if __name__ == "__main__":
    res = {"nb": nb, "rank": k, "host_cpus": os.cpu_count(), "threads": {}}
    for W in (1, 2, 4, 8, 16, 32):
        start, stop, counts = threading.Event(), threading.Event(), {}
        ths = [threading.Thread(target=persistent_worker, args=(w, start, stop, counts)) for w in range(W)]
        for th in ths:
            th.start()
        time.sleep(1.5 + 0.1 * W)                       # every worker has made its first calls
        t0 = time.perf_counter()
        start.set()
        time.sleep(3.0)
        stop.set()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        total = sum(counts.values())
        res["threads"][W] = {"tiles_per_s": total / dt, "ms_per_call": 1e3 * dt * W / max(total, 1)}
        print(f"HCORE_dgemm (libhcore_b200, host pointers) from {W:2d} worker threads: {total / dt:8.0f} tiles/s, "
              f"{1e3 * dt * W / max(total, 1):6.2f} ms per call", flush=True)
    try:
        import bench
        ref = bench.cpu_baseline(nb, k, 1e-8, 24)
        res["reference_all_cores"] = {"tiles_per_s": ref["value"], "cores": ref["cores"]}
        print(f"reference HCORE_dgemm, {ref['cores']} single-thread-BLAS processes: {ref['value']:.0f} tiles/s")
    except Exception as e:  # pragma: no cover
        print("reference not available:", e)
    print(json.dumps(res))
