"""Developer diagnostic (GPU box): run the TLR pipeline stage by stage and compare every
intermediate with the numpy model of the device algorithm (tests/device_model.py) and the
final result with the oracle.  Usage: python tools/gpu_stage_check.py [nb k ntiles]..."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import device_model as dm  # noqa: E402
from hcore_b200 import api  # noqa: E402
from hcore_b200.tilegen import make_lr_batch  # noqa: E402
from oracle import tlr_oracle as orc  # noqa: E402


def pick_width(maxrk):
    w = min(maxrk, 256)
    return 32 if w <= 32 else (64 if w <= 64 else (128 if w <= 128 else 256))


def run_case(nb, k, T, alpha=-1.0, beta=1.0, acc=1e-8, rk=0, verbose=True):
    print(f"=== case nb={nb} k={k} T={T} alpha={alpha} beta={beta} acc={acc} rk={rk}")
    b = make_lr_batch(T, nb, k)
    maxrk = 2 * k
    N = pick_width(maxrk)
    m_pad = max(128, (nb + 127) // 128 * 128)
    dev = torch.device("cuda:0")
    tens = {n: torch.from_numpy(getattr(b, n)).to(dev) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    ctx = api.Context(0)
    ok = True
    for stages in (1, 2, 3, 4):
        cu, cv, crk = tens["cu"].clone(), tens["cv"].clone(), tens["crk"].clone()
        status = torch.full((T,), -1, dtype=torch.int32, device=dev)
        ctx.debug_set_stages(stages)
        ctx.dgemm_batched(alpha, tens["au"], tens["av"], tens["ark"], tens["bu"], tens["bv"], tens["brk"], beta,
                          cu, cv, crk, rk=rk, maxrk=maxrk, acc=acc, status=status)
        torch.cuda.synchronize()
        for t in range(min(T, 2)):
            AU, AV, BU, BV = b.au[t].T, b.av[t].T, b.bu[t].T, b.bv[t].T
            CU, CV = b.cu[t][:k].T, b.cv[t][:k].T
            xu = np.hstack([beta * CU, alpha * AU])
            xv = np.hstack([CV, BU @ (AV.T @ BV).T])
            n = xu.shape[1]
            nblk, nspf = m_pad // 128, N // 8
            REC, RECV, LDT = 1152, 1024, 8
            if stages == 1:
                Gd = ctx.debug_fetch(7, t, m_pad * N).reshape(N, m_pad).T
                Gm = BU @ (AV.T @ BV).T
                e = np.abs(Gd[:nb, :Gm.shape[1]] - Gm).max()
                z = np.abs(Gd[nb:, :Gm.shape[1]]).max() if m_pad > nb else 0.0
                print(f"  [stage1 t={t}] G: max|dev-model|={e:.2e} padrows={z:.1e}")
                ok &= e < 1e-12 and z == 0
            if stages == 2:
                rr, cc = np.meshgrid(np.arange(128), np.arange(8), indexing="ij")
                hh = ((cc & 1) << 2) ^ (((cc >> 1) & 3) << 1)
                swz = (((cc << 6) + ((rr >> 1) ^ hh)) << 1) + (rr & 1)
                for w, nm, X in ((2, "Ru", xu), (3, "Rv", xv)):
                    f = dm.tsqr_factor(X)
                    R = ctx.debug_fetch(w, t, N * N).reshape(N, N).T
                    e = np.abs(R[:n, :n] - f["R"]).max()
                    low = np.abs(np.tril(R, -1)).max()
                    g = np.abs(R[:n, :n].T @ R[:n, :n] - X.T @ X).max()
                    recs = ctx.debug_fetch(w - 2, t, nblk * nspf * REC).reshape(nblk, nspf, REC)
                    ev = et = ett = 0.0
                    for j, (lo, hi) in enumerate(f["bounds"]):
                        for s_ in range(len(f["T"][j])):
                            V = recs[j, s_, :RECV][swz]                       # 128 x 8 explicit reflector panel
                            Vm = dm._vpanel(f, j, s_)
                            ev = max(ev, np.abs(V[:hi - lo, :Vm.shape[1]] - Vm).max())
                            Tm_ = recs[j, s_, RECV:RECV + 8 * LDT].reshape(8, LDT)[:, :8].T      # T[k][q] at k + 8 q
                            Tt_ = recs[j, s_, RECV + 8 * LDT:].reshape(8, LDT)[:, :8]            # T[k][q] at q + 8 k
                            pw = f["T"][j][s_].shape[0]
                            et = max(et, np.abs(Tm_[:pw, :pw] - f["T"][j][s_]).max())
                            ett = max(ett, np.abs(Tt_ - Tm_).max())
                    print(f"  [stage2 t={t}] {nm}: |R-model|={e:.2e} |tril(R)|={low:.1e} |RtR-XtX|={g:.2e} |V-model|={ev:.2e} "
                          f"|T-model|={et:.2e} |T-Tt|={ett:.1e}")
                    ok &= e < 1e-10 and g < 1e-11 and (et < 1e-10 or w == 3) and ett == 0   # V panel: tiny columns are rounding-sensitive
            if stages == 3:
                fu, fv = dm.tsqr_factor(xu), dm.tsqr_factor(xv)
                u, w_, sig, r, sweeps = dm.core_svd_model(fu["R"], fv["R"], rk, acc)
                ints = ctx.debug_fetch(9, t, 4)
                sg = ctx.debug_fetch(6, t, N)
                Yu = ctx.debug_fetch(4, t, N * N).reshape(N, N).T
                Yv = ctx.debug_fetch(5, t, N * N).reshape(N, N).T
                rd = int(ints[1])
                es = np.abs(sg[:n] - sig).max()
                # singular vectors are unique up to sign: compare projectors / products
                ru, rv = fu["R"], fv["R"]
                Tm = ru @ rv.T
                eprod = np.abs(Yu[:n, :rd] @ Yv[:n, :rd].T - u[:, :r] @ w_[:, :r].T).max() if rd == r else np.nan
                orth = np.abs(Yu[:n, :rd].T @ Yu[:n, :rd] - np.eye(rd)).max()
                print(f"  [stage3 t={t}] ncols={int(ints[0])} rank dev={rd} model={r} status={int(ints[2])} sweeps dev={int(ints[3])} "
                      f"model={sweeps} |sig-model|={es:.2e} |UW^T-model|={eprod:.2e} orth={orth:.1e} "
                      f"|UW^T - T_r|={np.abs(Yu[:n,:rd] @ Yv[:n,:rd].T - (Tm - (Tm - u[:, :r] @ w_[:, :r].T))).max():.2e}")
                ok &= rd == r and es < 1e-12
            if stages == 4:
                rnew = int(crk[t].item())
                cun, cvn = cu[t, :rnew].cpu().numpy().T, cv[t, :rnew].cpu().numpy().T
                D = b.dense_c(t, alpha, beta)
                err = np.linalg.norm(cun @ cvn.T - D)
                cu2, cv2, r2, sig2 = orc.tlr_gemm(AU, AV, BU, BV, CU, CV, alpha=alpha, beta=beta, rk=rk, maxrk=maxrk, acc=acc)
                err2 = np.linalg.norm(cu2 @ cv2.T - D)
                orth = np.abs(cun.T @ cun - np.eye(rnew)).max()
                print(f"  [stage4 t={t}] rank dev={rnew} oracle={r2} status={int(status[t].item())} err dev={err:.3e} oracle={err2:.3e} orthU={orth:.1e}")
                ok &= abs(rnew - r2) <= 1 and err <= max(err2 * 1.05, 1e-14 * np.linalg.norm(D)) + (acc if rnew != r2 else 0)
    ctx.close()
    print("  RESULT:", "PASS" if ok else "FAIL")
    return ok


if __name__ == "__main__":
    args = [int(x) for x in sys.argv[1:]]
    cases = [tuple(args[i:i + 3]) for i in range(0, len(args), 3)] or [(256, 8, 2), (500, 25, 2), (1024, 64, 2)]
    allok = True
    for (nb, k, T) in cases:
        allok &= run_case(nb, k, T)
    print("ALL:", "PASS" if allok else "FAIL")
    sys.exit(0 if allok else 1)
