"""GPU parity of the tile compression: HCORE_dcompress_batched_strided against the oracle's
restatement of the reference's compression procedure (dgesvd + rank rule, reference testing/testing_dgemm.c:119-236).
Floating point; gates: rank within +-1 of the oracle's (the update path's scan starts at index 1, so rank >= 1),
||A - U V^T||_F <= max(1.01 x oracle error, acc ||A||_F... see each assert), U orthonormal."""
import numpy as np
import pytest

from hcore_b200 import api
from hcore_b200.tilegen import _morton_points
from oracle import tlr_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available()
    c = api.Context(0)
    yield c
    c.close()


def _cov_tiles(T, nb, length, seed=3):
    """off-diagonal tiles of squared-exponential covariance matrices (the matrices HiCMA compresses)"""
    out = []
    for t in range(T):
        rng = np.random.default_rng(seed + t)
        pts = _morton_points(2 * nb, rng)
        d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
        out.append(np.exp(-(d / (length * (1 + 0.5 * (t % 3)))) ** 2)[nb:, :nb])
    return np.stack(out)


def _graded_tiles(T, nb, decay, seed=5):
    """orth * diag(decay^i) * orth^T: a slowly decaying spectrum with no gap at the threshold"""
    out = []
    for t in range(T):
        rng = np.random.default_rng(seed + t)
        q1 = np.linalg.qr(rng.random((nb, nb)) - 0.5)[0]; q2 = np.linalg.qr(rng.random((nb, nb)) - 0.5)[0]
        out.append((q1 * decay ** np.arange(nb)) @ q2.T)
    return np.stack(out)


def _run(ctx, A, acc, maxrk, power_iters=1):
    T, nb, _ = A.shape
    dev = torch.device("cuda:0")
    At = torch.from_numpy(np.ascontiguousarray(A.transpose(0, 2, 1))).to(dev)      # column-major tiles
    U = torch.zeros((T, maxrk, nb), dtype=torch.float64, device=dev); V = torch.zeros_like(U)
    rank = torch.full((T,), -1, dtype=torch.int32, device=dev); st = torch.full((T,), -1, dtype=torch.int32, device=dev)
    ctx.dcompress_batched(At, acc, maxrk, U, V, rank, status=st, power_iters=power_iters)
    torch.cuda.synchronize()
    assert torch.equal(At, torch.from_numpy(np.ascontiguousarray(A.transpose(0, 2, 1))).to(dev))          # A is read only
    return U.cpu().numpy(), V.cpu().numpy(), rank.cpu().numpy(), st.cpu().numpy()


@pytest.mark.parametrize("nb,maxrk,acc,kind", [(256, 64, 1e-8, "cov"), (512, 100, 1e-8, "cov"), (1024, 96, 1e-6, "cov"),
                                               (300, 80, 1e-6, "graded"), (128, 40, 1e-4, "graded")])
def test_compress_vs_oracle(ctx, nb, maxrk, acc, kind):
    T = 3
    A = _cov_tiles(T, nb, 0.3) if kind == "cov" else _graded_tiles(T, nb, 0.8)
    U, V, rank, st = _run(ctx, A, acc, maxrk)
    assert (st == 0).all()
    for t in range(T):
        u_o, v_o, r_o = orc.compress(A[t], acc, maxrk)
        r = int(rank[t])
        assert abs(r - max(r_o, 1)) <= 1, (r, r_o)
        Ut, Vt = U[t, :r].T, V[t, :r].T
        assert np.abs(Ut.T @ Ut - np.eye(r)).max() < 1e-12                        # U orthonormal
        e_dev, e_orc, na = np.linalg.norm(A[t] - Ut @ Vt.T), np.linalg.norm(A[t] - u_o @ v_o.T), np.linalg.norm(A[t])
        # within 1 % of the SVD truncation (or at rounding level), plus the one singular value a lower rank drops
        assert e_dev <= max(1.01 * e_orc, 64 * np.finfo(float).eps * na) + (acc if r < r_o else 0.0), (e_dev, e_orc)
        # the singular values carried by V are those of A
        s_dev = np.linalg.norm(Vt, axis=0)
        s_ref = np.linalg.svd(A[t], compute_uv=False)[:r]
        assert np.abs(s_dev - s_ref).max() <= 1e-10 * s_ref[0] + 0.02 * acc


def test_compress_rank_cap_and_feeds_the_update_path(ctx):
    """rank above maxrk -> truncated to maxrk like the reference's storage cap (testing_dgemm.c:202-205); the factors go
    straight into HCORE_dgemm_batched (same layout, same conventions)."""
    nb, maxrk, acc = 256, 24, 1e-10
    A = _graded_tiles(3, nb, 0.7)
    U, V, rank, st = _run(ctx, A, acc, maxrk)
    assert (st == 0).all() and (rank == maxrk).all()
    for t in range(3):
        u_o, v_o, r_o = orc.compress(A[t], acc, maxrk)
        assert r_o == maxrk
        e_dev, e_orc = np.linalg.norm(A[t] - U[t].T @ V[t]), np.linalg.norm(A[t] - u_o @ v_o.T)
        assert e_dev <= 1.01 * e_orc
    # C <- C - A B^T with all three tiles compressed on the device
    dev = torch.device("cuda:0")
    k = maxrk
    cu = torch.zeros((1, 2 * k, nb), dtype=torch.float64, device=dev); cv = torch.zeros_like(cu)
    cu[0, :k] = torch.from_numpy(U[2]).to(dev); cv[0, :k] = torch.from_numpy(V[2]).to(dev)
    rk = torch.full((1,), k, dtype=torch.int32, device=dev); crk = rk.clone()
    g = lambda x: torch.from_numpy(np.ascontiguousarray(x[None])).to(dev)
    ctx.dgemm_batched(-1.0, g(U[0]), g(V[0]), rk, g(U[1]), g(V[1]), rk, 1.0, cu, cv, crk, rk=0, maxrk=2 * k, acc=1e-8)
    torch.cuda.synchronize()
    r = int(crk[0].item())
    dense = U[2].T @ V[2] - (U[0].T @ V[0]) @ (U[1].T @ V[1]).T
    got = cu[0, :r].cpu().numpy().T @ cv[0, :r].cpu().numpy()
    assert np.linalg.norm(got - dense) <= 1e-8 * np.sqrt(2 * k)


def test_compress_exactly_low_rank_and_no_power_iteration(ctx):
    nb, r0 = 384, 17
    rng = np.random.default_rng(1)
    A = np.stack([(rng.random((nb, r0)) - 0.5) @ (rng.random((r0, nb)) - 0.5) for _ in range(2)])
    for q in (0, 1):
        U, V, rank, st = _run(ctx, A, 1e-9, 64, power_iters=q)
        assert (st == 0).all() and (rank == r0).all()
        for t in range(2):
            assert np.linalg.norm(A[t] - U[t, :r0].T @ V[t, :r0]) <= 1e-12 * np.linalg.norm(A[t])
