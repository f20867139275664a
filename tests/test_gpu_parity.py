"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against (1) the reference's own known-answer tests, (2) committed outputs of the compiled
reference, (3) the oracle on seeded inputs, (4) size-independent properties at full size.
Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from hcore_b200 import api
from hcore_b200.tilegen import make_lr_batch, make_dense_batch, make_covariance_batch
from oracle import tlr_oracle as orc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    c = api.Context(0)
    yield c
    c.close()


def _inf_norm(x):
    return np.abs(x).sum(axis=1).max()


def _to_dev(b):
    dev = torch.device("cuda:0")
    return {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev)
            for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}


def _run_batch(ctx, b, alpha, beta, rk, maxrk, acc):
    t = _to_dev(b)
    status = torch.full((b.ntiles,), -7, dtype=torch.int32, device="cuda:0")
    ctx.dgemm_batched(alpha, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], beta, t["cu"], t["cv"], t["crk"],
                      rk=rk, maxrk=maxrk, acc=acc, status=status)
    torch.cuda.synchronize()
    return t["cu"].cpu().numpy(), t["cv"].cpu().numpy(), t["crk"].cpu().numpy(), status.cpu().numpy()


# ---- (1) the reference's own known-answer tests, through the legacy drop-in symbols ----------------

def test_ctest_dgemm_legacy_symbol(golden_dir):
    """reference testing/testing_dgemm.c:250-392 replayed on HCORE_dgemm of libhcore_b200.so."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_dgemm.npz"))
    n, acc = int(g["n"]), float(g["acc"])
    ra, rb_, rc = (int(x) for x in g["ranks"])
    smr = 6                                               # storage_maxrank: CU/CV buffers are n x 6 (:259,:287-290)
    CU = np.zeros((smr, n)); CU[:rc] = g["CU"][:rc]
    CV = np.zeros((smr, n)); CV[:rc] = g["CV"][:rc]
    Ark, Brk, Crk = np.array([float(ra)]), np.array([float(rb_)]), np.array([float(rc)])
    fl = api.FlopCounter()
    api.hcore_dgemm(api.HCORE_NoTrans, api.HCORE_Trans, n, n, 1.0, np.ascontiguousarray(g["AU"]),
                    np.ascontiguousarray(g["AV"]), Ark, n, np.ascontiguousarray(g["BU"]), np.ascontiguousarray(g["BV"]),
                    Brk, n, 1.0, CU, CV, Crk, n, 0, n, acc, None, fl)
    r = int(Crk[0])
    assert r == int(g["out_rank"]) == 6                   # "Crank: 4 -> 6"
    rnorm = _inf_norm(g["ref_c"] - CU[:r].T @ CV[:r])
    assert rnorm < 10 * acc                               # the reference's assert (:377-378)
    assert abs(rnorm - float(g["rnorm"])) < 1e-10         # 1.476392e-04
    assert fl.update == int(g["flops"])                   # same flop accounting as the reference


def test_ctest_sgemm_legacy_symbol(golden_dir):
    """reference testing/testing_sgemm.c replayed on HCORE_sgemm of libhcore_b200.so (float in/out,
    FP64 arithmetic inside): same rank, same assert, residual no worse than the reference's."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_sgemm.npz"))
    n, acc = int(g["n"]), float(g["acc"])
    ra, rb_, rc = (int(x) for x in g["ranks"])
    CU = np.zeros((n, n), np.float32); CU[:rc] = g["CU"][:rc]
    CV = np.zeros((n, n), np.float32); CV[:rc] = g["CV"][:rc]
    Ark, Brk, Crk = (np.array([float(x)], np.float32) for x in (ra, rb_, rc))
    fl = api.FlopCounter()
    api.hcore_sgemm(api.HCORE_NoTrans, api.HCORE_Trans, n, n, 1.0, np.ascontiguousarray(g["AU"]),
                    np.ascontiguousarray(g["AV"]), Ark, n, np.ascontiguousarray(g["BU"]), np.ascontiguousarray(g["BV"]),
                    Brk, n, 1.0, CU, CV, Crk, n, 0, n, acc, None, fl)
    r = int(Crk[0])
    assert r == int(g["out_rank"]) == 6
    rnorm = _inf_norm(g["ref_c"].astype(np.float64) - CU[:r].T.astype(np.float64) @ CV[:r].astype(np.float64))
    assert rnorm < 10 * acc                               # the reference's assert (testing_sgemm.c:377-378)
    assert rnorm <= float(g["rnorm"]) * 1.05
    assert fl.update == int(g["flops"])


@pytest.mark.parametrize("nb,k,acc", [(256, 16, 1e-4), (1024, 64, 1e-4), (500, 25, 1e-3)])
def test_sgemm_batched_vs_oracle(ctx, nb, k, acc):
    """HCORE_sgemm_batched (config 4, float leg) against the float32 oracle (sgeqrf/sgesvd/sormqr)."""
    T = 3
    b = make_lr_batch(T, nb, k, base_seed=41, dtype=np.float32, decades=4.0)
    dev = torch.device("cuda:0")
    t = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev)
         for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    status = torch.full((T,), -7, dtype=torch.int32, device=dev)
    ctx.sgemm_batched(-1.0, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, t["cu"], t["cv"], t["crk"],
                      rk=0, maxrk=2 * k, acc=acc, status=status)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    cu, cv, crk = t["cu"].cpu().numpy(), t["cv"].cpu().numpy(), t["crk"].cpu().numpy()
    assert cu.dtype == np.float32
    for i in range(T):
        cu_o, cv_o, r_o, _ = orc.tlr_gemm_colmajor(b.au[i], b.av[i], b.bu[i], b.bv[i], b.cu[i], b.cv[i], k,
                                                   alpha=np.float32(-1.0), beta=np.float32(1.0), rk=0, maxrk=2 * k, acc=acc)
        r = int(crk[i])
        assert abs(r - r_o) <= 1
        dense = b.dense_c(i, -1.0, 1.0).astype(np.float64)
        e_o = np.linalg.norm(cu_o.T.astype(np.float64) @ cv_o.astype(np.float64) - dense)
        e_d = np.linalg.norm(cu[i, :r].T.astype(np.float64) @ cv[i, :r].astype(np.float64) - dense)
        # FP64 inside: no worse than the float reference (plus float rounding of the output factors)
        assert e_d <= 1.01 * e_o + 8 * np.finfo(np.float32).eps * np.linalg.norm(dense) + (acc if r != r_o else 0.0)


@pytest.mark.parametrize("nb,k,acc,rk", [(96, 8, 1e-8, 0), (256, 16, 1e-8, 0), (200, 12, 1e-8, 5), (1024, 64, 1e-8, 0),
                                         (384, 100, 1e-8, 0), (512, 128, 1e-8, 0), (300, 80, 1e-8, 150),   # stack width 256
                                         (257, 9, 1e-8, 0)])                # odd nb: 16-byte access path of the blocked QR
def test_zgemm_batched_vs_oracle(ctx, nb, k, acc, rk):
    """HCORE_zgemm_batched (config 4, complex leg) against the complex oracle (zgeqrf/zgesvd/zunmqr,
    plain-transpose tiles, reference src/compute/hcore_zgemm.c:165,189,309)."""
    T = 2
    alpha, beta = -1.0 + 0.25j, 1.0 - 0.5j
    b = make_lr_batch(T, nb, k, base_seed=51, dtype=np.complex128)
    dev = torch.device("cuda:0")
    t = {n: torch.from_numpy(np.ascontiguousarray(getattr(b, n))).to(dev)
         for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    status = torch.full((T,), -7, dtype=torch.int32, device=dev)
    ctx.zgemm_batched(alpha, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], beta, t["cu"], t["cv"], t["crk"],
                      rk=rk, maxrk=2 * k, acc=acc, status=status)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    cu, cv, crk = t["cu"].cpu().numpy(), t["cv"].cpu().numpy(), t["crk"].cpu().numpy()
    for i in range(T):
        cu_o, cv_o, r_o, _ = orc.tlr_gemm_colmajor(b.au[i], b.av[i], b.bu[i], b.bv[i], b.cu[i], b.cv[i], k,
                                                   alpha=alpha, beta=beta, rk=rk, maxrk=2 * k, acc=acc)
        r = int(crk[i])
        assert abs(r - r_o) <= (0 if rk else 1)
        dense = b.dense_c(i, alpha, beta)
        e_o = np.linalg.norm(cu_o.T @ cv_o - dense)
        e_d = np.linalg.norm(cu[i, :r].T @ cv[i, :r] - dense)
        assert e_d <= max(1.01 * e_o, 256 * EPS * np.linalg.norm(dense)) + (acc if r != r_o else 0.0)
        if rk == 0:
            U = cu[i, :r]
            assert np.abs(U.conj() @ U.T - np.eye(r)).max() < 1e-12     # CU has orthonormal columns


def test_zgemm_legacy_symbol(ctx):
    """HCORE_zgemm drop-in symbol (host arrays, double _Complex scalars by value)."""
    nb, k = 64, 6
    alpha, beta = 0.5 - 1.0j, 1.0 + 0.0j
    b = make_lr_batch(1, nb, k, base_seed=77, dtype=np.complex128)
    CU, CV = b.cu[0].copy(), b.cv[0].copy()
    Ark, Brk, Crk = np.array([float(k)]), np.array([float(k)]), np.array([float(k)])
    fl = api.FlopCounter()
    api.hcore_zgemm(api.HCORE_NoTrans, api.HCORE_Trans, nb, nb, alpha, b.au[0], b.av[0], Ark, nb, b.bu[0], b.bv[0], Brk, nb,
                    beta, CU, CV, Crk, nb, 0, 2 * k, 1e-8, None, fl)
    r = int(Crk[0])
    cu_o, cv_o, r_o, _ = orc.tlr_gemm_colmajor(b.au[0], b.av[0], b.bu[0], b.bv[0], b.cu[0], b.cv[0], k, alpha=alpha,
                                               beta=beta, rk=0, maxrk=2 * k, acc=1e-8)
    assert abs(r - r_o) <= 1
    dense = b.dense_c(0, alpha, beta)
    assert np.linalg.norm(CU[:r].T @ CV[:r] - dense) <= 1.01 * np.linalg.norm(cu_o.T @ cv_o - dense) + 1e-8
    assert fl.update == api.gemm_flops(nb, k, k, k, r)


def test_ctest_dgemmcd_legacy_symbol(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_ctest_dgemmcd.npz"))
    CD = np.array(g["CD"], copy=True)
    api.hcore_dgemmcd(api.HCORE_NoTrans, api.HCORE_Trans, 10, -1.0, np.ascontiguousarray(g["Au"]),
                      np.ascontiguousarray(g["Av"]), 7, 10, np.ascontiguousarray(g["Bu"]), np.ascontiguousarray(g["Bv"]),
                      6, 10, 1.0, CD, 10, 7)
    dense = g["CD"].T - (g["Au"].T @ g["Av"]) @ (g["Bu"].T @ g["Bv"]).T
    assert _inf_norm(dense - CD.T) < 1000 * EPS           # reference testing_dgemmcd.c:194-205
    assert np.abs(CD - g["out"]).max() < 1000 * EPS


def test_ctest_dsyrk_legacy_symbol(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_ctest_dsyrk.npz"))
    CD = np.array(g["CD"], copy=True)
    fl = api.FlopCounter()
    api.hcore_dsyrk(api.HCORE_Lower, api.HCORE_NoTrans, 10, 7, -1.0, np.ascontiguousarray(g["AU"]), 10,
                    np.ascontiguousarray(g["AV"]), 10, 1.0, CD, 10, None, fl)
    Am = g["AU"].T @ g["AV"]
    dense = np.tril(g["CD"].T - Am @ Am.T)
    assert _inf_norm(dense - np.tril(CD.T)) < 1000 * EPS  # reference testing_dsyrk.c:170-181
    assert np.abs(np.triu(CD.T, 1)).max() == 0.0          # strict upper triangle zeroed (hcore_dsyrk.c:174-181)
    assert np.abs(CD - g["out"]).max() < 1000 * EPS
    assert fl.syrk == int(g["flops"])


# ---- (2) committed outputs of the compiled reference --------------------------------------------------

@pytest.mark.parametrize("name,kind", [("nb64_k8", None), ("nb96_ragged", "ragged"), ("nb64_fixed", None)])
def test_small_batches_vs_reference_outputs(ctx, golden_dir, name, kind):
    g = np.load(os.path.join(golden_dir, "ref_tlr_small.npz"))
    nb, k, rk, T = (int(x) for x in g[name + "_meta"])
    acc, alpha, beta = (float(x) for x in g[name + "_scal"])
    ranks = [tuple(int(v) for v in row) for row in g[name + "_ranks"]] if kind == "ragged" else None
    b = make_lr_batch(T, nb, k, base_seed=99, ranks=ranks)
    cu, cv, crk, st = _run_batch(ctx, b, alpha, beta, rk, 2 * k, acc)
    assert (st == 0).all()
    for t in range(T):
        r_ref = int(g[f"{name}_t{t}_rank"])
        r = int(crk[t])
        assert abs(r - r_ref) <= 1
        ref = g[f"{name}_t{t}_cu"].T @ g[f"{name}_t{t}_cv"]
        dense = b.dense_c(t, alpha, beta)
        e_ref = np.linalg.norm(ref - dense)
        e_dev = np.linalg.norm(cu[t, :r].T @ cv[t, :r] - dense)
        if r == r_ref:
            assert np.abs(cu[t, :r].T @ cv[t, :r] - ref).max() < 1e-12
        assert e_dev <= max(1.01 * e_ref, 64 * EPS * np.linalg.norm(dense)) + (acc if r != r_ref else 0.0)
        if rk == 0:
            assert np.abs(cu[t, :r] @ cu[t, :r].T - np.eye(r)).max() < 1e-13   # U factor stays orthonormal


@pytest.mark.parametrize("name,nb,k,T", [("nb500_k25", 500, 25, 3), ("nb1024_k64", 1024, 64, 2)])
def test_large_tiles_vs_reference_outputs(ctx, golden_dir, name, nb, k, T):
    """BASELINE.json configs[0] (nb=500, rank 25) and configs[1] (nb=1024, rank 64), acc 1e-8."""
    g = np.load(os.path.join(golden_dir, "ref_tlr_large.npz"))
    b = make_lr_batch(T, nb, k)
    cu, cv, crk, st = _run_batch(ctx, b, -1.0, 1.0, 0, 2 * k, 1e-8)
    assert (st == 0).all()
    for t in range(T):
        r, r_ref = int(crk[t]), int(g[name + "_ranks"][t])
        assert abs(r - r_ref) <= 1
        dense = b.dense_c(t, -1.0, 1.0)
        err = np.linalg.norm(cu[t, :r].T @ cv[t, :r] - dense)
        assert err <= max(1e-8 * np.linalg.norm(dense), 1.01 * float(g[name + "_err"][t])) + (1e-8 if r != r_ref else 0)
        assert err / np.linalg.norm(dense) <= 1e-8       # north star: relative error at or below the accuracy


# ---- (3) oracle on seeded inputs: shapes, scalars, modes -------------------------------------------------

@pytest.mark.parametrize("nb,k,alpha,beta,rk,acc", [
    (128, 16, -1.0, 1.0, 0, 1e-8), (130, 9, 1.0, 1.0, 0, 1e-6), (257, 17, 0.3, -2.0, 0, 1e-10),
    (1000, 40, -1.0, 1.0, 0, 1e-8), (384, 64, -1.0, 1.0, 0, 1e-8), (200, 12, -1.0, 1.0, 7, 1e-8),
    (200, 12, -1.0, 1.0, 100, 1e-8), (64, 3, -1.0, 1.0, 0, 1e-3), (2048, 32, -1.0, 1.0, 0, 1e-8),
    # ranks up to 128: stack width 256, the core matrix of the SVD kernel lives in L2 instead of shared memory
    (512, 128, -1.0, 1.0, 0, 1e-8), (300, 100, -1.0, 1.0, 0, 1e-8), (1024, 96, -1.0, 1.0, 150, 1e-8),
    (4096, 24, -1.0, 1.0, 0, 1e-8),           # 32 row blocks per panel
    # the sweep corners of BASELINE.json configs[4] whose throughput profiles/ops_sweep_*.json quotes
    (2048, 64, -1.0, 1.0, 0, 1e-8), (2048, 128, -1.0, 1.0, 0, 1e-8), (1024, 128, -1.0, 1.0, 0, 1e-8),
    (512, 64, -1.0, 1.0, 0, 1e-8), (512, 16, -1.0, 1.0, 0, 1e-8), (1024, 16, -1.0, 1.0, 0, 1e-8),
    (1024, 32, -1.0, 1.0, 0, 1e-8), (512, 32, -1.0, 1.0, 0, 1e-8),
    # degenerate scalars: half (or all) of the stacked panel is zero, the pivoted QR stops early on exact zeros
    (256, 24, 0.0, 1.0, 0, 1e-8), (256, 24, -1.0, 0.0, 0, 1e-8), (192, 10, 0.0, 0.0, 0, 1e-8), (1024, 64, 0.0, 1.0, 0, 1e-8)])
def test_vs_oracle(ctx, nb, k, alpha, beta, rk, acc):
    T = 3
    b = make_lr_batch(T, nb, k, base_seed=7 + nb)
    cu, cv, crk, st = _run_batch(ctx, b, alpha, beta, rk, 2 * k, acc)
    assert (st == 0).all()
    for t in range(T):
        cu_o, cv_o, r_o, sig = orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k,
                                                     alpha=alpha, beta=beta, rk=rk, maxrk=2 * k, acc=acc)
        r = int(crk[t])
        assert abs(r - r_o) <= (0 if rk else 1)
        dense = b.dense_c(t, alpha, beta)
        e_o = np.linalg.norm(cu_o.T @ cv_o - dense)
        e_d = np.linalg.norm(cu[t, :r].T @ cv[t, :r] - dense)
        assert e_d <= max(1.01 * e_o, 256 * EPS * np.linalg.norm(dense)) + (acc if r != r_o else 0.0)


@pytest.mark.parametrize("nb,maxrk,acc", [(256, 128, 1e-8), (500, 128, 1e-6)])
def test_covariance_tiles_vs_oracle(ctx, nb, maxrk, acc):
    """The realistic regime: A, B, C cut from one exponential-kernel covariance matrix (ragged ranks,
    update and C share most of their column space, rank after recompression close to the ranks in)."""
    T = 4
    b = make_covariance_batch(T, nb, maxrk, acc=acc)
    cu, cv, crk, st = _run_batch(ctx, b, -1.0, 1.0, 0, maxrk, acc)
    assert (st == 0).all()
    for t in range(T):
        a_, b_, c_ = int(b.ark[t]), int(b.brk[t]), int(b.crk[t])
        cu_o, cv_o, r_o, sig = orc.tlr_gemm_colmajor(b.au[t, :a_], b.av[t, :a_], b.bu[t, :b_], b.bv[t, :b_], b.cu[t], b.cv[t],
                                                     c_, alpha=-1.0, beta=1.0, rk=0, maxrk=maxrk, acc=acc)
        r = int(crk[t])
        assert abs(r - r_o) <= 1
        assert r < a_ + c_                                   # the stack really was rank deficient
        dense = b.dense_c(t, -1.0, 1.0)
        e_o = np.linalg.norm(cu_o.T @ cv_o - dense)
        e_d = np.linalg.norm(cu[t, :r].T @ cv[t, :r] - dense)
        assert e_d <= max(1.01 * e_o, 256 * EPS * np.linalg.norm(dense)) + (acc if r != r_o else 0.0)


def test_empty_batch_is_a_no_op(ctx):
    z = torch.zeros((0, 8, 64), dtype=torch.float64, device="cuda:0")
    zi = torch.zeros((0,), dtype=torch.int32, device="cuda:0")
    ctx.dgemm_batched(-1.0, z, z, zi, z, z, zi, 1.0, torch.zeros((0, 16, 64), dtype=torch.float64, device="cuda:0"),
                      torch.zeros((0, 16, 64), dtype=torch.float64, device="cuda:0"), zi, rk=0, maxrk=16, acc=1e-8)
    torch.cuda.synchronize()


def test_ragged_ranks_and_rank_deficient_stack(ctx):
    nb, k = 192, 16
    ranks = [(16, 16, 16), (1, 1, 1), (3, 16, 9), (16, 2, 1), (7, 7, 7)]
    b = make_lr_batch(len(ranks), nb, k, base_seed=21, ranks=ranks)
    b.au[4, :7] = b.cu[4, :7]                            # A and C share their column space exactly
    cu, cv, crk, st = _run_batch(ctx, b, -1.0, 1.0, 0, 2 * k, 1e-8)
    assert (st == 0).all()
    for t, (a, bb, c) in enumerate(ranks):
        cu_o, cv_o, r_o, _ = orc.tlr_gemm(b.au[t, :a].T, b.av[t, :a].T, b.bu[t, :bb].T, b.bv[t, :bb].T, b.cu[t, :c].T,
                                          b.cv[t, :c].T, alpha=-1.0, beta=1.0, rk=0, maxrk=2 * k, acc=1e-8)
        r = int(crk[t])
        assert abs(r - r_o) <= 1
        dense = b.dense_c(t, -1.0, 1.0)
        assert np.linalg.norm(cu[t, :r].T @ cv[t, :r] - dense) <= 1.01 * np.linalg.norm(cu_o @ cv_o.T - dense) + 2e-8
    assert int(crk[4]) <= 8                               # rank-deficient stack is really recompressed


def test_per_tile_status_instead_of_exit(ctx):
    """Errors that make the reference exit(-1) (hcore_dgemm.c:743-746, :924-927) become per-tile status."""
    nb, k = 128, 8
    ranks = [(8, 8, 8), (8, 8, 0), (8, 8, 8), (0, 8, 8)]
    b = make_lr_batch(4, nb, k, base_seed=2, ranks=[(8, 8, 8)] * 4)
    b.ark[:] = [r[0] for r in ranks]; b.crk[:] = [r[2] for r in ranks]
    before = b.cu.copy()
    cu, cv, crk, st = _run_batch(ctx, b, -1.0, 1.0, 0, 2 * k, 1e-8)
    assert st.tolist() == [0, 1, 0, 1]
    assert np.array_equal(cu[1], before[1]) and np.array_equal(cu[3], before[3])    # failed tiles untouched
    b2 = make_lr_batch(2, nb, k, base_seed=2)
    cu, cv, crk, st = _run_batch(ctx, b2, -1.0, 1.0, 0, 2 * k - 1, 1e-8)            # maxrk < Ark + Crk
    assert st.tolist() == [2, 2] and crk.tolist() == [8, 8]


# ---- (4) properties at the benchmark size ------------------------------------------------------------------

def test_full_size_properties_and_shard_invariance(ctx):
    """nb=1024, rank 64, acc 1e-8 (BASELINE.json configs[1]) on 96 tiles: every tile meets the
    accuracy bound, U stays orthonormal, ranks sit where the reference's do (92-94), and the
    per-tile results do not depend on how the batch is cut (tiles are independent)."""
    nb, k, T, acc = 1024, 64, 96, 1e-8
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    import bench
    d = bench.make_device_batch(torch, T, nb, k, dev, seed=5)
    bench.restore_c(d, k)
    ctx.dgemm_batched(-1.0, d["au"], d["av"], d["ark"], d["bu"], d["bv"], d["brk"], 1.0, d["cu"], d["cv"], d["crk"],
                      rk=0, maxrk=2 * k, acc=acc, status=d["status"])
    torch.cuda.synchronize()
    assert int(d["status"].abs().sum().item()) == 0
    ranks = d["crk"].cpu().numpy()
    assert ranks.min() >= 90 and ranks.max() <= 96
    full_cu, full_cv = d["cu"].clone(), d["cv"].clone()
    # EVERY tile: accuracy bound and orthonormal U factor
    err, dn = _batch_errors(torch, d["au"], d["av"], d["bu"], d["bv"], d["cu0"], d["cv0"], d["cu"], d["cv"],
                            d["crk"].long(), -1.0, 1.0)
    assert (err <= acc * np.sqrt(2 * k)).all()            # reference guarantee: ||err||_F <= acc*sqrt(#dropped)
    assert (err / dn <= 1.25 * acc).all() and (err / dn <= acc).float().mean().item() >= 0.95   # relative: at or near acc
    for t in range(T):
        r = int(ranks[t])
        U = d["cu"][t, :r]
        assert (U @ U.T - torch.eye(r, dtype=torch.float64, device=dev)).abs().max().item() < 1e-13
    # same batch cut in two unequal shards and in small chunks: bit-identical per-tile output
    for cut, chunk in ((37, 0), (64, 5)):
        ctx.set_chunk(chunk)
        bench.restore_c(d, k)
        for lo, hi in ((0, cut), (cut, T)):
            ctx.dgemm_batched(-1.0, d["au"][lo:hi], d["av"][lo:hi], d["ark"][lo:hi], d["bu"][lo:hi], d["bv"][lo:hi],
                              d["brk"][lo:hi], 1.0, d["cu"][lo:hi], d["cv"][lo:hi], d["crk"][lo:hi], rk=0,
                              maxrk=2 * k, acc=acc, status=d["status"][lo:hi])
        torch.cuda.synchronize()
        assert np.array_equal(d["crk"].cpu().numpy(), ranks)
        for t in range(T):
            r = int(ranks[t])
            assert torch.equal(d["cu"][t, :r], full_cu[t, :r]) and torch.equal(d["cv"][t, :r], full_cv[t, :r])
    ctx.set_chunk(0)


def _batch_errors(torch, au, av, bu, bv, cu0, cv0, cu, cv, ranks, alpha, beta):
    """||CU CV^T - (alpha A B^T + beta C)||_F and ||dense||_F per tile, on the GPU (verification only)."""
    T, kc, nb = cu.shape
    col = torch.arange(kc, device=cu.device)[None, :, None]
    mask = (col < ranks.to(cu.device)[:, None, None]).to(cu.dtype)
    approx = torch.bmm((cu * mask).transpose(1, 2), cv * mask)
    A = torch.bmm(au.transpose(1, 2), av)
    B = torch.bmm(bu.transpose(1, 2), bv)
    dense = alpha * torch.bmm(A, B.transpose(1, 2)) + beta * torch.bmm(cu0.transpose(1, 2), cv0)
    return torch.linalg.norm((approx - dense).flatten(1), dim=1), torch.linalg.norm(dense.flatten(1), dim=1)


def test_same_inputs_as_the_reference_256_tiles(ctx):
    """North star: "ranks equal to the reference within +-1 ON THE SAME INPUTS" and "meets the accuracy bound
    on EVERY tile": 256 generated tiles at the benchmark shape (nb=1024, rank 64, acc 1e-8) go through the
    compiled reference (oracle/_ref, reference src/compute/hcore_dgemm.c:703-1012; the numpy oracle when the
    compiled reference is absent) and through HCORE_dgemm_batched; the truncation rule compared is
    hcore_dgemm.c:444-468."""
    from oracle import ref_binding as rb
    nb, k, T, acc = 1024, 64, 256, 1e-8
    b = make_lr_batch(T, nb, k, base_seed=4242)
    cu_d, cv_d, crk_d, st = _run_batch(ctx, b, -1.0, 1.0, 0, 2 * k, acc)
    assert (st == 0).all()
    use_ref = rb.available()
    if use_ref:
        rb.set_blas_threads(1)
    cu_r = np.zeros_like(b.cu); cv_r = np.zeros_like(b.cv); crk_r = np.zeros(T, np.int64)
    for t in range(T):
        if use_ref:
            cu2, cv2, r, _ = rb.hcore_gemm(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0, beta=1.0,
                                           rk=0, maxrk=2 * k, acc=acc)
        else:
            cu2, cv2, r, _ = orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0,
                                                   beta=1.0, rk=0, maxrk=2 * k, acc=acc)
        cu_r[t, :r], cv_r[t, :r], crk_r[t] = cu2[:r], cv2[:r], r
    dev = torch.device("cuda:0")
    g = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    au, av, bu, bv, cu0, cv0 = (g(x) for x in (b.au, b.av, b.bu, b.bv, b.cu[:, :k], b.cv[:, :k]))
    e_dev, dn = _batch_errors(torch, au, av, bu, bv, cu0, cv0, g(cu_d), g(cv_d), torch.from_numpy(crk_d.astype(np.int64)), -1.0, 1.0)
    e_ref, _ = _batch_errors(torch, au, av, bu, bv, cu0, cv0, g(cu_r), g(cv_r), torch.from_numpy(crk_r), -1.0, 1.0)
    e_dev, e_ref, dn = e_dev.cpu().numpy(), e_ref.cpu().numpy(), dn.cpu().numpy()
    dr = crk_d.astype(np.int64) - crk_r
    assert np.abs(dr).max() <= 1, (dr.min(), dr.max())
    # error <= max(acc * ||C||, 1.01 x the reference's own error); a tile whose rank is one lower than the
    # reference's additionally carries the singular value it dropped (< acc by the truncation rule)
    bound = np.maximum(acc * dn, 1.01 * e_ref) + np.where(dr < 0, acc, 0.0)
    assert (e_dev <= bound).all(), float((e_dev / bound).max())
    # relative error at or below the requested accuracy on every tile -- except where the reference's own result is
    # above it too (its guarantee is absolute: ||err||_F <= acc * sqrt(#dropped), hcore_dgemm.c:453-464)
    assert ((e_dev / dn <= acc) | (e_ref / dn > acc)).all()
    print(f"same-input parity: {T} tiles, rank diffs {np.bincount(dr + 1, minlength=3).tolist()} (-1/0/+1), "
          f"max err ratio {float((e_dev / np.maximum(e_ref, 1e-300)).max()):.4f}, reference = {'compiled' if use_ref else 'numpy oracle'}")


def test_host_buffer_path_matches_device_path(ctx):
    nb, k, T = 256, 16, 700                               # > one 512-tile staging slice
    b = make_lr_batch(8, nb, k, base_seed=31)
    rep = lambda x: np.ascontiguousarray(np.tile(x, (T // 8 + 1,) + (1,) * (x.ndim - 1))[:T])
    import copy
    bb = copy.copy(b)
    for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk"):
        setattr(bb, n, rep(getattr(b, n)))
    cu_d, cv_d, crk_d, st_d = _run_batch(ctx, bb, -1.0, 1.0, 0, 2 * k, 1e-8)
    cu_h, cv_h, crk_h = bb.cu.copy(), bb.cv.copy(), bb.crk.copy()
    st_h = np.full(T, -1, np.int32)
    ctx.dgemm_batched_host(-1.0, bb.au, bb.av, bb.ark, bb.bu, bb.bv, bb.brk, 1.0, cu_h, cv_h, crk_h, rk=0, maxrk=2 * k,
                           acc=1e-8, status=st_h)
    assert (st_h == 0).all() and np.array_equal(crk_h, crk_d)
    for t in range(T):
        r = int(crk_d[t])
        assert np.array_equal(cu_h[t, :r], cu_d[t, :r]) and np.array_equal(cv_h[t, :r], cv_d[t, :r])
        assert not cu_h[t, r:2 * k].any() and not cv_h[t, r:2 * k].any()      # columns past the new rank come back zeroed


@pytest.mark.parametrize("dt,acc", [(np.float32, 1e-4), (np.complex128, 1e-8)])
def test_host_buffer_path_single_and_complex(ctx, dt, acc):
    """HCORE_sgemm_batched_host / HCORE_zgemm_batched_host: the same three-stream pipeline, bit-identical to the device call."""
    nb, k, T = 192, 12, 400
    b = make_lr_batch(4, nb, k, base_seed=33, dtype=dt, decades=4.0 if dt == np.float32 else 8.0)
    rep = lambda x: np.ascontiguousarray(np.tile(x, (T // 4 + 1,) + (1,) * (x.ndim - 1))[:T])
    h = {n: rep(getattr(b, n)) for n in ("au", "av", "bu", "bv", "cu", "cv", "ark", "brk", "crk")}
    t = {n: torch.from_numpy(h[n]).to("cuda:0") for n in h}
    st_d = torch.full((T,), -7, dtype=torch.int32, device="cuda:0")
    call = ctx.sgemm_batched if dt == np.float32 else ctx.zgemm_batched
    al, be = (-1.0, 1.0) if dt == np.float32 else (-1.0 + 0.5j, 1.0 - 0.25j)
    call(al, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], be, t["cu"], t["cv"], t["crk"], rk=0, maxrk=2 * k, acc=acc,
         status=st_d)
    torch.cuda.synchronize()
    st_h = np.full(T, -1, np.int32)
    ctx.dgemm_batched_host(al, h["au"], h["av"], h["ark"], h["bu"], h["bv"], h["brk"], be, h["cu"], h["cv"], h["crk"], rk=0,
                           maxrk=2 * k, acc=acc, status=st_h)
    assert (st_h == 0).all() and np.array_equal(h["crk"], t["crk"].cpu().numpy())
    cu_d, cv_d = t["cu"].cpu().numpy(), t["cv"].cpu().numpy()
    for i in range(T):
        r = int(h["crk"][i])
        assert np.array_equal(h["cu"][i, :r], cu_d[i, :r]) and np.array_equal(h["cv"][i, :r], cv_d[i, :r])


# ---- dense-C updates (BASELINE.json configs[2]) ----------------------------------------------------------------

@pytest.mark.parametrize("nb,ka,kb", [(1024, 32, 32), (500, 25, 13), (130, 7, 6), (1024, 64, 96), (257, 9, 5), (129, 3, 40)])
def test_dgemmcd_dsyrk_batched_vs_oracle(ctx, nb, ka, kb):
    T = 3
    b = make_lr_batch(T, nb, max(ka, kb), base_seed=17, ranks=[(ka, kb, 1)] * T)
    t = _to_dev(b)
    cd = make_dense_batch(T, nb)
    c1 = torch.from_numpy(cd).to("cuda:0")
    ctx.dgemmcd_batched(t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], c1)
    c2 = torch.from_numpy(cd).to("cuda:0")
    ctx.dsyrk_batched(-1.0, t["au"], t["av"], t["ark"], c2)
    torch.cuda.synchronize()
    for i in range(T):
        ref = orc.gemmcd(b.au[i, :ka].T, b.av[i, :ka].T, b.bu[i, :kb].T, b.bv[i, :kb].T, cd[i].T)
        assert np.abs(c1[i].cpu().numpy().T - ref).max() < 1000 * EPS * max(1.0, np.abs(ref).max())
        ref = orc.syrk(b.au[i, :ka].T, b.av[i, :ka].T, cd[i].T, -1.0)
        out = c2[i].cpu().numpy().T
        assert np.abs(out - ref).max() < 1000 * EPS * max(1.0, np.abs(ref).max())
        assert np.abs(np.triu(out, 1)).max() == 0.0


# ---- (5) split-phase API: rank first, factors into exact-size arrays afterwards ------------------------------

def test_split_phase_batched_equals_fused(ctx):
    """HCORE_dgemm_batched_qr_svd + HCORE_dgemm_batched_ormqr (reference HCORE_dgemm_qr_svd / _ormqr,
    src/compute/hcore_dgemm.c:1312-1699) give the fused call's factors bit for bit, written into freshly
    allocated arrays of exactly max(new rank) columns; phase 1 leaves C untouched."""
    nb, k, T = 320, 24, 5
    b = make_lr_batch(T, nb, k, base_seed=77)
    cu_f, cv_f, crk_f, st_f = _run_batch(ctx, b, -1.0, 1.0, 0, 2 * k, 1e-8)
    t = _to_dev(b)
    newrank = torch.zeros(T, dtype=torch.int32, device="cuda:0")
    status = torch.full((T,), -7, dtype=torch.int32, device="cuda:0")
    cu_before, crk_before = t["cu"].clone(), t["crk"].clone()
    ctx.dgemm_batched_qr_svd(-1.0, t["au"], t["av"], t["ark"], t["bu"], t["bv"], t["brk"], 1.0, t["cu"], t["cv"], t["crk"],
                             newrank, rk=0, maxrk=2 * k, acc=1e-8, status=status)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    assert torch.equal(t["cu"], cu_before) and torch.equal(t["crk"], crk_before)
    assert np.array_equal(newrank.cpu().numpy(), crk_f)
    rmax = int(newrank.max().item())
    cu_new = torch.zeros((T, rmax, nb), dtype=torch.float64, device="cuda:0")
    cv_new = torch.zeros_like(cu_new)
    ctx.dgemm_batched_ormqr(cu_new, cv_new, t["crk"])
    torch.cuda.synchronize()
    assert np.array_equal(t["crk"].cpu().numpy(), crk_f)
    for i in range(T):
        r = int(crk_f[i])
        assert np.array_equal(cu_new[i, :r].cpu().numpy(), cu_f[i, :r])
        assert np.array_equal(cv_new[i, :r].cpu().numpy(), cv_f[i, :r])
    # a second half without a first half is refused
    with pytest.raises(Exception):
        ctx.dgemm_batched_ormqr(cu_new, cv_new, t["crk"])


def test_split_phase_legacy_symbols():
    """HCORE_dgemm_qr_svd / HCORE_dgemm_ormqr with the reference's prototypes (host pointers, int ranks):
    same factors and flop count as the one-call HCORE_dgemm; the caller re-allocates C in between."""
    import ctypes as C
    from hcore_b200 import _lib
    L = _lib.load()
    nb, k = 200, 12
    b = make_lr_batch(1, nb, k, base_seed=5)
    au, av, bu, bv = (np.ascontiguousarray(x[0]) for x in (b.au, b.av, b.bu, b.bv))
    cu = np.zeros((2 * k, nb)); cv = np.zeros((2 * k, nb)); cu[:k] = b.cu[0][:k]; cv[:k] = b.cv[0][:k]
    ref_cu, ref_cv, crk = cu.copy(), cv.copy(), np.array([float(k)])
    fl1 = api.FlopCounter()
    api.hcore_dgemm(111, 112, nb, nb, -1.0, au, av, np.array([float(k)]), nb, bu, bv, np.array([float(k)]), nb, 1.0,
                    ref_cu, ref_cv, crk, nb, 0, 2 * k, 1e-8, flops=fl1)
    r_ref = int(crk[0])
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    P = lambda x: x.ctypes.data_as(dp)
    ark, brk, ck = (C.c_int * 1)(k), (C.c_int * 1)(k), (C.c_int * 1)(k)
    outs_p = [dp() for _ in range(11)]; outs_i = [C.c_int() for _ in range(8)]
    fl2 = api.FlopCounter()
    o = outs_p; n = outs_i
    L.HCORE_dgemm_qr_svd(111, 112, nb, nb, -1.0, P(au), P(av), ark, nb, P(bu), P(bv), brk, nb, 1.0, P(cu), P(cv), ck, nb,
                         0, 2 * k, 2 * k, 1e-8, None, C.byref(fl2),
                         C.byref(o[0]), C.byref(o[1]), C.byref(o[2]), C.byref(n[0]), C.byref(n[1]), C.byref(o[3]), C.byref(n[2]),
                         C.byref(o[4]), C.byref(n[3]), C.byref(o[5]), C.byref(n[4]), C.byref(o[6]), C.byref(n[5]), C.byref(o[7]),
                         C.byref(n[6]), C.byref(o[8]), C.byref(o[9]), C.byref(n[7]), C.byref(o[10]))
    rnew = n[1].value
    assert rnew == r_ref and ck[0] == k and n[0].value == 2 * k
    # the caller re-allocates C at the revealed rank (leading dimension unchanged) and finishes the update
    cu2 = np.zeros((rnew, nb)); cv2 = np.zeros((rnew, nb))
    L.HCORE_dgemm_ormqr(111, 112, nb, nb, -1.0, P(au), P(av), ark, nb, 1.0, P(cu2), P(cv2), ck, nb, 0, 2 * k, 2 * k, 1e-8,
                        None, C.byref(fl2), o[1], o[2], n[0].value, rnew, o[3], n[2].value, o[4], n[3].value, o[5], n[4].value,
                        o[6], n[5].value, o[7], n[6].value, P(cu2), o[9], n[7].value, P(cv2))
    assert ck[0] == rnew
    assert np.array_equal(cu2, ref_cu[:rnew]) and np.array_equal(cv2, ref_cv[:rnew])
    assert fl2.update == fl1.update


# ---- (6) B-dense variant ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("nb,k", [(256, 16), (200, 12), (1024, 64), (129, 5)])
def test_b_dense_batched_vs_oracle(ctx, nb, k):
    """C <- alpha (AU AV^T) B^T + beta C with a DENSE B tile (reference HCORE_dgemm_qr_svd_b_dense + _ormqr,
    src/compute/hcore_dgemm.c:1022-1303).  Oracle: the same update with B = B * I^T (BU = B, BV = identity)."""
    T = 2
    b = make_lr_batch(T, nb, k, base_seed=31 + nb)
    rng = np.random.default_rng(5)
    Bd = (rng.random((T, nb, nb)) - 0.5) / np.sqrt(nb)            # [t][col][row]: column-major nb x nb
    t = _to_dev(b)
    Bt = torch.from_numpy(Bd).to("cuda:0")
    status = torch.full((T,), -7, dtype=torch.int32, device="cuda:0")
    ctx.dgemm_batched_b_dense(-1.0, t["au"], t["av"], t["ark"], Bt, 1.0, t["cu"], t["cv"], t["crk"], rk=0, maxrk=2 * k,
                              acc=1e-8, status=status)
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    cu, cv, crk = t["cu"].cpu().numpy(), t["cv"].cpu().numpy(), t["crk"].cpu().numpy()
    eye = np.eye(nb)
    for i in range(T):
        Bm = Bd[i].T                                              # math orientation
        cu_o, cv_o, r_o, _ = orc.tlr_gemm(b.au[i].T, b.av[i].T, Bm, eye, b.cu[i][:k].T, b.cv[i][:k].T, alpha=-1.0, beta=1.0,
                                          rk=0, maxrk=2 * k, acc=1e-8)
        r = int(crk[i])
        assert abs(r - r_o) <= 1
        dense = b.cu[i][:k].T @ b.cv[i][:k] - (b.au[i].T @ b.av[i]) @ Bm.T
        e_o = np.linalg.norm(cu_o @ cv_o.T - dense)
        e_d = np.linalg.norm(cu[i, :r].T @ cv[i, :r] - dense)
        assert e_d <= max(1.01 * e_o, 256 * EPS * np.linalg.norm(dense)) + (1e-8 if r != r_o else 0.0)


def test_b_dense_legacy_split_symbols():
    """HCORE_dgemm_qr_svd_b_dense + HCORE_dgemm_ormqr (host pointers, reference prototypes) equal the batched call."""
    import ctypes as C
    from hcore_b200 import _lib
    L = _lib.load()
    nb, k = 160, 10
    b = make_lr_batch(1, nb, k, base_seed=9)
    rng = np.random.default_rng(2)
    Bd = np.ascontiguousarray((rng.random((nb, nb)) - 0.5) / np.sqrt(nb))
    c = api.Context(0)
    t = _to_dev(b)
    Bt = torch.from_numpy(Bd[None]).to("cuda:0")
    c.dgemm_batched_b_dense(-1.0, t["au"], t["av"], t["ark"], Bt, 1.0, t["cu"], t["cv"], t["crk"], rk=0, maxrk=2 * k, acc=1e-8)
    torch.cuda.synchronize()
    r_ref = int(t["crk"][0].item())
    ref_cu, ref_cv = t["cu"][0, :r_ref].cpu().numpy(), t["cv"][0, :r_ref].cpu().numpy()
    c.close()
    au, av = np.ascontiguousarray(b.au[0]), np.ascontiguousarray(b.av[0])
    cu = np.zeros((2 * k, nb)); cv = np.zeros((2 * k, nb)); cu[:k] = b.cu[0][:k]; cv[:k] = b.cv[0][:k]
    dp = C.POINTER(C.c_double)
    P = lambda x: x.ctypes.data_as(dp)
    ark, ck = (C.c_int * 1)(k), (C.c_int * 1)(k)
    o = [dp() for _ in range(11)]; n = [C.c_int() for _ in range(8)]
    fl = api.FlopCounter()
    L.HCORE_dgemm_qr_svd_b_dense(111, 112, nb, nb, -1.0, P(au), P(av), ark, nb, P(Bd), nb, 1.0, P(cu), P(cv), ck, nb,
                                 0, 2 * k, 2 * k, 1e-8, None, C.byref(fl),
                                 C.byref(o[0]), C.byref(o[1]), C.byref(o[2]), C.byref(n[0]), C.byref(n[1]), C.byref(o[3]), C.byref(n[2]),
                                 C.byref(o[4]), C.byref(n[3]), C.byref(o[5]), C.byref(n[4]), C.byref(o[6]), C.byref(n[5]), C.byref(o[7]),
                                 C.byref(n[6]), C.byref(o[8]), C.byref(o[9]), C.byref(n[7]), C.byref(o[10]))
    rnew = n[1].value
    assert rnew == r_ref
    cu2 = np.zeros((rnew, nb)); cv2 = np.zeros((rnew, nb))
    L.HCORE_dgemm_ormqr(111, 112, nb, nb, -1.0, P(au), P(av), ark, nb, 1.0, P(cu2), P(cv2), ck, nb, 0, 2 * k, 2 * k, 1e-8,
                        None, C.byref(fl), o[1], o[2], n[0].value, rnew, o[3], n[2].value, o[4], n[3].value, o[5], n[4].value,
                        o[6], n[5].value, o[7], n[6].value, P(cu2), o[9], n[7].value, P(cv2))
    assert ck[0] == rnew and fl.update > 0
    assert np.array_equal(cu2, ref_cu) and np.array_equal(cv2, ref_cv)
