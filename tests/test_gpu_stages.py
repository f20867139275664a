"""Stage-level GPU parity (rows a2-a6 of SURVEY.md section 8): every intermediate of the batched update against
the LAPACK call of the reference it replaces, on the same stacked panels.

  stage 1  G = BU (AV^T BV)^T          <-> the two cblas_dgemm of __qrb   (reference src/compute/hcore_dgemm.c:175,:199)
  stage 2  R_u, R_v of the TSQR        <-> LAPACKE_dgeqrf of __qra / __qrb (:123, :216)   (R^T R, |R| up to row signs)
  stage 3  sigma, rank, U_r, W         <-> LAPACKE_dgesvd of R_u R_v^T + truncation (:384-:468)
  stage 4  CU = Q_u [U_r;0], CV = Q_v [W;0] <-> dormqr / dge_trans of __newu / __newv (:543, :655-:673)

The intermediates are read back with hcore_b200_debug_fetch after running only the first k kernels of the pipeline
(hcore_b200_debug_set_stages).  Nothing here reads /root/reference."""
import numpy as np
import pytest
import scipy.linalg as sla

from hcore_b200 import api
from hcore_b200.tilegen import make_lr_batch, make_covariance_batch

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _width(maxrk):
    w = min(maxrk, 256)
    return 32 if w <= 32 else (64 if w <= 64 else (128 if w <= 128 else 256))


def _run_stages(ctx, b, stages, alpha, beta, rk, maxrk, acc):
    dev = torch.device("cuda:0")
    t = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev)
         for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    status = torch.full((b.ntiles,), -1, dtype=torch.int32, device=dev)
    ctx.debug_set_stages(stages)
    try:
        ctx.dgemm_batched(alpha, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], beta, t["cu"], t["cv"], t["crk"],
                          rk=rk, maxrk=maxrk, acc=acc, status=status)
        torch.cuda.synchronize()
    finally:
        ctx.debug_set_stages(4)
    return t, status


def _stacks(b, t, alpha, beta):
    a, bb, c = int(b.ark[t]), int(b.brk[t]), int(b.crk[t])
    AU, AV, BU, BV = b.au[t, :a].T, b.av[t, :a].T, b.bu[t, :bb].T, b.bv[t, :bb].T
    CU, CV = b.cu[t, :c].T, b.cv[t, :c].T
    G = BU @ (AV.T @ BV).T
    return np.hstack([beta * CU, alpha * AU]), np.hstack([CV, G]), G


CASES = [(256, 8, "lr"), (500, 25, "lr"), (1024, 64, "lr"), (384, 100, "lr"), (512, 128, "cov")]


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available()
    c = api.Context(0)
    yield c
    c.close()


def _batch(nb, k, kind):
    if kind == "cov":
        return make_covariance_batch(2, nb, k, acc=1e-8), k
    return make_lr_batch(2, nb, k, base_seed=11 + nb), 2 * k


@pytest.mark.parametrize("nb,k,kind", CASES)
def test_stage1_stacking_products(ctx, nb, k, kind):
    b, maxrk = _batch(nb, k, kind)
    N, m_pad = _width(maxrk), max(128, (nb + 127) // 128 * 128)
    _run_stages(ctx, b, 1, -1.0, 1.0, 0, maxrk, 1e-8)
    for t in range(b.ntiles):
        _, _, G = _stacks(b, t, -1.0, 1.0)
        Gd = ctx.debug_fetch(7, t, m_pad * N).reshape(N, m_pad).T
        scale = max(np.abs(G).max(), 1e-300)
        assert np.abs(Gd[:nb, :G.shape[1]] - G).max() <= 64 * np.finfo(float).eps * scale * np.sqrt(nb)
        if m_pad > nb:
            assert np.abs(Gd[nb:, :G.shape[1]]).max() == 0.0          # padding rows stay exact zeros


@pytest.mark.parametrize("nb,k,kind", CASES)
def test_stage2_tsqr_against_dgeqrf(ctx, nb, k, kind):
    b, maxrk = _batch(nb, k, kind)
    N = _width(maxrk)
    _run_stages(ctx, b, 2, -1.0, 1.0, 0, maxrk, 1e-8)
    for t in range(b.ntiles):
        xu, xv, _ = _stacks(b, t, -1.0, 1.0)
        n = xu.shape[1]
        for what, X in ((2, xu), (3, xv)):
            R = ctx.debug_fetch(what, t, N * N).reshape(N, N).T[:n, :n]
            Rl = np.triu(sla.lapack.dgeqrf(X)[0][:n, :n])
            nx = np.linalg.norm(X, 2)
            assert np.abs(np.tril(R, -1)).max() == 0.0
            # R^T R = X^T X to rounding: R is a Cholesky factor of the same Gram matrix as dgeqrf's
            assert np.abs(R.T @ R - X.T @ X).max() <= 256 * np.finfo(float).eps * nx * nx
            assert np.abs(R.T @ R - Rl.T @ Rl).max() <= 256 * np.finfo(float).eps * nx * nx
            if kind == "lr" and what == 2:
                # well-conditioned stack: R is unique up to the sign of each row
                assert np.abs(np.abs(R) - np.abs(Rl)).max() <= 1e-10 * nx
            # the singular values of R are those of the panel (what the SVD stage consumes)
            sd, sl = np.linalg.svd(R, compute_uv=False), np.linalg.svd(Rl, compute_uv=False)
            assert np.abs(sd - sl).max() <= 64 * np.finfo(float).eps * nx * np.sqrt(n)


@pytest.mark.parametrize("rk", [0, 7])
@pytest.mark.parametrize("nb,k,kind", CASES)
def test_stage3_core_svd_against_dgesvd(ctx, nb, k, kind, rk):
    acc = 1e-8
    b, maxrk = _batch(nb, k, kind)
    N = _width(maxrk)
    _, status = _run_stages(ctx, b, 3, -1.0, 1.0, rk, maxrk, acc)
    assert (status.cpu().numpy() == 0).all()
    for t in range(b.ntiles):
        xu, xv, _ = _stacks(b, t, -1.0, 1.0)
        n = xu.shape[1]
        Ru = np.triu(sla.lapack.dgeqrf(xu)[0][:n, :n])
        Rv = np.triu(sla.lapack.dgeqrf(xv)[0][:n, :n])
        T = Ru @ Rv.T                                         # reference :384
        u, s, vt, info = sla.lapack.dgesvd(T)                 # reference :403
        assert info == 0
        if rk:
            r_ref = min(rk, n)                                # :448-452
        else:
            below = np.nonzero(s[1:] < acc)[0]                # :453-464, scan from index 1
            r_ref = int(below[0]) + 1 if below.size else n
        ints = ctx.debug_fetch(9, t, 4)
        r = int(ints[1])
        assert abs(r - r_ref) <= (0 if rk else 1)
        sg = ctx.debug_fetch(6, t, N)[:n]
        # singular values down to the truncation threshold (the pivoted QR drops a block of norm < acc / 100)
        keep = s >= (acc if not rk else 1e-13 * s[0])
        assert np.abs(sg[keep] - s[keep]).max() <= (0.02 * acc if not rk else 1e-12 * s[0]) + 1e-13 * s[0]
        Yu = ctx.debug_fetch(4, t, N * N).reshape(N, N).T[:n, :r]
        Yv = ctx.debug_fetch(5, t, N * N).reshape(N, N).T[:n, :r]
        assert np.abs(Yu.T @ Yu - np.eye(r)).max() < 1e-13    # U_r orthonormal
        # U_r W^T is a rank-r approximation of the device's own T = R_u R_v^T as good as dgesvd's truncation
        Rud = ctx.debug_fetch(2, t, N * N).reshape(N, N).T[:n, :n]
        Rvd = ctx.debug_fetch(3, t, N * N).reshape(N, N).T[:n, :n]
        Td = Rud @ Rvd.T
        e_dev = np.linalg.norm(Yu @ Yv.T - Td)
        e_opt = np.sqrt((s[r:] ** 2).sum())
        assert e_dev <= 1.01 * e_opt + 1e-13 * s[0] + (0.02 * acc if not rk else 0.0)


@pytest.mark.parametrize("nb,k,kind", CASES)
def test_stage4_reflector_application(ctx, nb, k, kind):
    """CU = Q_u [U_r; 0] and CV = Q_v [W; 0]: CU orthonormal, CU^T (stacked U panel) = U_r^T R_u, and the product equals
    the stacked update restricted to the kept subspace."""
    acc = 1e-8
    b, maxrk = _batch(nb, k, kind)
    N = _width(maxrk)
    # the small factors U_r, W and the R factors as the SVD stage leaves them (the apply kernel updates Y in place)
    _, status = _run_stages(ctx, b, 3, -1.0, 1.0, 0, maxrk, acc)
    assert (status.cpu().numpy() == 0).all()
    small = []
    for t in range(b.ntiles):
        r = int(ctx.debug_fetch(9, t, 4)[1])
        small.append([ctx.debug_fetch(w, t, N * N).reshape(N, N).T.copy() for w in (4, 5, 2, 3)] + [r])
    t_, status = _run_stages(ctx, b, 4, -1.0, 1.0, 0, maxrk, acc)
    assert (status.cpu().numpy() == 0).all()
    cu, cv, crk = t_["cu"].cpu().numpy(), t_["cv"].cpu().numpy(), t_["crk"].cpu().numpy()
    for t in range(b.ntiles):
        xu, xv, _ = _stacks(b, t, -1.0, 1.0)
        n, r = xu.shape[1], int(crk[t])
        assert r == small[t][4]
        CU, CV = cu[t, :r].T, cv[t, :r].T
        assert np.abs(CU.T @ CU - np.eye(r)).max() < 1e-13
        Yu, Yv, Rud, Rvd = small[t][0][:n, :r], small[t][1][:n, :r], small[t][2][:n, :n], small[t][3][:n, :n]
        nx = np.linalg.norm(xu, 2) * np.linalg.norm(xv, 2)
        # Q_u^T X_u = R_u  =>  CU^T X_u = U_r^T R_u ; likewise CV^T X_v = W^T R_v
        assert np.abs(CU.T @ xu - Yu.T @ Rud).max() <= 1e-11 * np.linalg.norm(xu, 2)
        assert np.abs(CV.T @ xv - Yv.T @ Rvd).max() <= 1e-11 * np.linalg.norm(Yv, 2) * np.linalg.norm(xv, 2)   # entries of W^T R_v are of the order ||W|| ||X_v||
        # and the product is the stacked update restricted to the kept subspace: error below the dropped singular values
        assert np.linalg.norm(CU @ CV.T - xu @ xv.T) <= max(acc * np.sqrt(n), 1e-13 * nx)
