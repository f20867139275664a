"""Phase cycles of the complex core SVD kernel (instrumented build).  GPU box."""
import ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HCORE_B200_LIB", os.path.join(ROOT, "hcore_b200", "lib", "libhcore_b200_prof.so"))
from hcore_b200 import api, _lib
from hcore_b200.tilegen import make_lr_batch
dev = torch.device("cuda:0")
ctx = api.Context(0)
L = _lib.load()
L.hcore_b200_prof_read.argtypes = [C.c_void_p, C.c_int, C.c_int]
T, nb, k = 148, 1024, 64
b = make_lr_batch(4, nb, k, base_seed=5, dtype=np.complex128)
t = {n: torch.from_numpy(np.ascontiguousarray(np.concatenate([getattr(b, n)] * (T // 4)))).to(dev) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
buf = (C.c_ulonglong * 64)()
st = torch.zeros(T, dtype=torch.int32, device=dev)
for _ in range(2):
    crk = t["crk"].clone(); cu = t["cu"].clone(); cv = t["cv"].clone()
    L.hcore_b200_prof_read(buf, 64, 1)
    ctx.zgemm_batched(-1.0, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, cu, cv, crk, rk=0, maxrk=2 * k, acc=1e-8, status=st)
    torch.cuda.synchronize()
    L.hcore_b200_prof_read(buf, 64, 0)
v = list(buf); n = max(v[39], 1)
names = {30: "T^H product", 31: "QRCP", 32: "R^H", 33: "Jacobi", 34: "sort / U_r", 35: "W products"}
print("tiles", n, "kq", v[37] / n, "sweeps", v[38] / n, "status", int(st.abs().sum()))
for s_, nm in names.items():
    print(f"  {nm:16s} {v[s_] / n:12,.0f} cycles per tile")
if v[49]:
    print("stack kernel: copy %.0f, P %.0f, G %.0f cycles per tile" % (v[46] / v[49], v[47] / v[49], v[48] / v[49]))
if v[55]:
    print("QR kernel per panel: block load %.0f, in-block %.0f, T factor %.0f, trailing %.0f, write-back %.0f cycles" % tuple(v[i] / v[55] for i in (50, 51, 52, 53, 54)))
    print("  trailing update per panel (warp 0): load %.0f, step 1 %.0f, reduce + barriers %.0f, step 2 %.0f, store + rest %.0f cycles" % (tuple(v[i] / v[55] for i in (56, 57, 58, 59)) + (v[53] / v[55],)))
