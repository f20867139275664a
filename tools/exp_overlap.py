"""Experiment (GPU box): do two independent chunk pipelines on two streams fill each other's latency bubbles?

Two contexts (two workspaces) process the two halves of one batch on two CUDA streams; the second stream starts with
a short first call so that the two pipelines run out of phase (one in its SVD stage while the other is in its QR /
apply stages).  Compared with one context on one stream over the whole batch.  HCORE_B200_LIB selects the build
(SVD kernel register cap)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from hcore_b200 import api

nb, k = 1024, 64
dev = torch.device("cuda:0")
T = int(os.environ.get("T", 9472))
d = bench.make_device_batch(torch, T, nb, k, dev, seed=1)
ctxs = [api.Context(0), api.Context(0)]
streams = [torch.cuda.Stream(priority=int(os.environ.get("PRIO0", 0))), torch.cuda.Stream(priority=int(os.environ.get("PRIO1", 0)))]


def run(ctx, lo, hi, stream):
    if hi <= lo:
        return
    s = slice(lo, hi)
    ctx.dgemm_batched(-1.0, d["au"][s], d["av"][s], d["ark"][s], d["bu"][s], d["bv"][s], d["brk"][s], 1.0, d["cu"][s], d["cv"][s],
                      d["crk"][s], rk=0, maxrk=2 * k, acc=1e-8, status=d["status"][s], stream=stream)


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        bench.restore_c(d, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts


def single(chunk):
    ctxs[0].set_chunk(chunk)
    return timed(lambda: run(ctxs[0], 0, T, None))


def dual(chunk, offset):
    for c in ctxs:
        c.set_chunk(chunk)
    half = T // 2

    def fn():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        # interleave the host-side enqueue chunk by chunk so that neither stream's launches sit behind the other's
        a, b = 0, half
        first = True
        while a < half or b < T:
            if a < half:
                run(ctxs[0], a, min(a + chunk, half), streams[0]); a += chunk
            if b < T:
                n = offset if (first and offset > 0) else chunk
                run(ctxs[1], b, min(b + n, T), streams[1]); b += n
                first = False
        for s in streams:
            cur.wait_stream(s)
    return timed(fn)


print("lib", os.environ.get("HCORE_B200_LIB", "default"), "T", T, flush=True)
for chunk in (2368, 1184):
    t, ts = single(chunk)
    print("single chunk", chunk, "tiles/s", round(T / t * 1e3), [round(x, 1) for x in ts], flush=True)
for chunk in (1184, 2368):
    for off in (0, chunk // 2):
        t, ts = dual(chunk, off)
        print("dual chunk", chunk, "offset", off, "tiles/s", round(T / t * 1e3), [round(x, 1) for x in ts], flush=True)
st = d["status"].cpu()
print("status ok:", int((st == 0).sum()), "of", T)
