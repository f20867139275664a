"""Raw host<->device copy ceiling of this box, measured the way HCORE_dgemm_batched_host moves data: pinned host
buffers, slices of one SVD wave, H2D and D2H on separate streams at the same time, every rank at once.

  python tools/pcie_ceiling.py [out.json]                                  # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         tools/pcie_ceiling.py [out.json]                                   # N GPUs at once

The e2e roofline of bench.py divides its achieved H2D+D2H GB/s by the `both` figure measured here."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

nb, k, tiles = 1024, 64, 148
in_bytes = tiles * 6 * k * nb * 8            # six rank-k factors per tile go in
out_bytes = tiles * 2 * 93 * nb * 8          # two rank-93 factors per tile come back
nslices = 8
h_in = torch.empty(nslices * in_bytes, dtype=torch.uint8).pin_memory()
h_out = torch.empty(nslices * out_bytes, dtype=torch.uint8).pin_memory()
h_in.fill_(1); h_out.fill_(1)
d_in = torch.empty(2 * in_bytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(2 * out_bytes, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(do_in, do_out, reps=3):
    best = 1e9
    for _ in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        for s in range(nslices):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in[(s & 1) * in_bytes:(s & 1) * in_bytes + in_bytes].copy_(h_in[s * in_bytes:(s + 1) * in_bytes], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out[s * out_bytes:(s + 1) * out_bytes].copy_(d_out[(s & 1) * out_bytes:(s & 1) * out_bytes + out_bytes], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


res = {"n_gpus": world, "slice_tiles": tiles, "h2d_bytes_per_slice": in_bytes, "d2h_bytes_per_slice": out_bytes, "slices": nslices}
t = run(True, False); res["h2d_only_gbs"] = world * nslices * in_bytes / t / 1e9
t = run(False, True); res["d2h_only_gbs"] = world * nslices * out_bytes / t / 1e9
t = run(True, True)
res["both_h2d_gbs"] = world * nslices * in_bytes / t / 1e9
res["both_d2h_gbs"] = world * nslices * out_bytes / t / 1e9
res["both_gbs"] = res["both_h2d_gbs"] + res["both_d2h_gbs"]
res["tiles_per_s_ceiling"] = world * nslices * tiles / t
res["host_cpus"] = os.cpu_count()
if rank == 0:
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
