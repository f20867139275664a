"""Phase cycles of dense_stream_kernel (instrumented build, thread 0 of every CTA).  GPU box."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HCORE_B200_LIB", os.path.join(ROOT, "hcore_b200", "lib", "libhcore_b200_prof.so"))
from hcore_b200 import api, _lib
dev = torch.device("cuda:0")
ctx = api.Context(0)
L = _lib.load()
L.hcore_b200_prof_read.argtypes = [C.c_void_p, C.c_int, C.c_int]
T, nb, k = 1024, 1024, 32
g = torch.Generator(device=dev); g.manual_seed(1)
f = lambda: torch.rand((T, k, nb), dtype=torch.float64, device=dev, generator=g) - 0.5
au, av, bu, bv = f(), f(), f(), f()
cd = torch.rand((T, nb, nb), dtype=torch.float64, device=dev, generator=g)
rk = torch.full((T,), k, dtype=torch.int32, device=dev)
buf = (C.c_ulonglong * 64)()
for _ in range(2):
    ctx.dgemmcd_batched(au, av, rk, bu, bv, rk, cd)
    torch.cuda.synchronize()
    L.hcore_b200_prof_read(buf, 64, 1)
v = list(buf); n = max(v[63], 1)
names = {56: "loop top / sync tail", 57: "Y fetch + product", 58: "wait for the CD block", 59: "update", 60: "Y stage + fence + barrier", 61: "store + wait_group", 62: "produce"}
print("items", n)
for s, nm in names.items():
    print(f"  {nm:32s} {v[s] / n:9.0f} cycles per item")
