"""CPU tests: the oracle (numpy/LAPACK restatement) against the reference's own known-answer
tests and against outputs of the compiled, unmodified reference (tests/golden, produced by
tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from hcore_b200.tilegen import make_lr_batch
from oracle import ref_binding as rb
from oracle import tlr_oracle as orc
import device_model as dm


def _inf_norm(x):
    return np.abs(x).sum(axis=1).max()


def test_reference_ctest_dgemm(golden_dir):
    """reference testing/testing_dgemm.c: ranks 4/4/4 -> 6 and ||C_ref - CU CV^T||_inf < 10*acc (:377-378)."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_dgemm.npz"))
    ra, rb_, rc = g["ranks"]
    cu, cv, r, _ = orc.tlr_gemm(g["AU"][:ra].T, g["AV"][:ra].T, g["BU"][:rb_].T, g["BV"][:rb_].T, g["CU"][:rc].T,
                                g["CV"][:rc].T, alpha=1.0, beta=1.0, rk=0, maxrk=int(g["n"]), acc=float(g["acc"]))
    assert r == int(g["out_rank"]) == 6
    rn = _inf_norm(g["ref_c"] - cu @ cv.T)
    assert rn < 10 * float(g["acc"])
    assert abs(rn - float(g["rnorm"])) < 1e-12          # golden value 1.476392e-04
    assert abs(float(g["rnorm"]) - 1.476392e-04) < 1e-9
    np.testing.assert_allclose(cu.T, g["out_cu"], atol=1e-12)
    np.testing.assert_allclose(cv.T, g["out_cv"], atol=1e-12)
    assert orc.flop_model(10, ra, rb_, rc, r) == int(g["flops"])


def test_reference_ctest_sgemm(golden_dir):
    """reference testing/testing_sgemm.c (float clone, storage_maxrank 10): ranks 4/5/5 -> 6,
    ||C_ref - CU CV^T||_inf = 8.181651e-05 < 10*acc."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_sgemm.npz"))
    ra, rb_, rc = g["ranks"]
    assert (int(ra), int(rb_), int(rc)) == (4, 5, 5)
    assert g["AU"].dtype == np.float32
    cu, cv, r, _ = orc.tlr_gemm(g["AU"][:ra].T, g["AV"][:ra].T, g["BU"][:rb_].T, g["BV"][:rb_].T, g["CU"][:rc].T,
                                g["CV"][:rc].T, alpha=np.float32(1.0), beta=np.float32(1.0), rk=0, maxrk=int(g["n"]),
                                acc=float(g["acc"]))
    assert cu.dtype == np.float32
    assert r == int(g["out_rank"]) == 6
    rn = _inf_norm(g["ref_c"] - cu @ cv.T)
    assert rn < 10 * float(g["acc"])
    assert abs(float(g["rnorm"]) - 8.181651e-05) < 1e-9
    assert abs(rn - float(g["rnorm"])) < 1e-5            # float LAPACK: last digits depend on the BLAS
    np.testing.assert_allclose(cu @ cv.T, g["out_cu"].T @ g["out_cv"], atol=2e-5)


def test_reference_ctest_dgemmcd(golden_dir):
    """reference testing/testing_dgemmcd.c:194-205: residual < 1000 eps."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_dgemmcd.npz"))
    out = orc.gemmcd(g["Au"].T, g["Av"].T, g["Bu"].T, g["Bv"].T, g["CD"].T)
    assert np.abs(out - g["out"].T).max() < 1000 * np.finfo(float).eps


def test_reference_ctest_dsyrk(golden_dir):
    """reference testing/testing_dsyrk.c:170-181: lower triangle compared, residual < 1000 eps;
    the reference also zeroes the strict upper triangle (src/compute/hcore_dsyrk.c:174-181)."""
    g = np.load(os.path.join(golden_dir, "ref_ctest_dsyrk.npz"))
    out = orc.syrk(g["AU"].T, g["AV"].T, g["CD"].T, -1.0)
    assert np.abs(out - g["out"].T).max() < 1000 * np.finfo(float).eps
    assert np.abs(np.triu(out, 1)).max() == 0.0


CASES = [("nb64_k8", None), ("nb96_ragged", "ragged"), ("nb64_fixed", None)]


@pytest.mark.parametrize("name,kind", CASES)
def test_oracle_vs_compiled_reference_small(golden_dir, name, kind):
    g = np.load(os.path.join(golden_dir, "ref_tlr_small.npz"))
    nb, k, rk, T = (int(x) for x in g[name + "_meta"])
    acc, alpha, beta = (float(x) for x in g[name + "_scal"])
    ranks = [tuple(int(v) for v in row) for row in g[name + "_ranks"]] if kind == "ragged" else None
    b = make_lr_batch(T, nb, k, base_seed=99, ranks=ranks)
    for t in range(T):
        a, bb, c = int(b.ark[t]), int(b.brk[t]), int(b.crk[t])
        cu, cv, r, _ = orc.tlr_gemm(b.au[t, :a].T, b.av[t, :a].T, b.bu[t, :bb].T, b.bv[t, :bb].T, b.cu[t, :c].T,
                                    b.cv[t, :c].T, alpha=alpha, beta=beta, rk=rk, maxrk=2 * k, acc=acc)
        assert r == int(g[f"{name}_t{t}_rank"])
        np.testing.assert_allclose(cu @ cv.T, g[f"{name}_t{t}_cu"].T @ g[f"{name}_t{t}_cv"], atol=1e-12)
        assert orc.flop_model(nb, a, bb, c, r) == int(g[f"{name}_t{t}_flops"])


@pytest.mark.parametrize("name", ["z_nb64_k6", "z_nb96_fixed"])
def test_oracle_vs_compiled_reference_complex(golden_dir, name):
    """The complex restatement against outputs of the compiled reference HCORE_zgemm (the
    reference ships no zgemm test: this fixture is what pins the complex oracle)."""
    g = np.load(os.path.join(golden_dir, "ref_tlr_complex.npz"))
    nb, k, rk, T = (int(x.real) for x in g[name + "_meta"])
    acc, alpha, beta = g[name + "_scal"]
    b = make_lr_batch(T, nb, k, base_seed=51, dtype=np.complex128)
    for t in range(T):
        cu, cv, r, _ = orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=alpha,
                                             beta=beta, rk=rk, maxrk=2 * k, acc=float(acc.real))
        assert r == int(g[f"{name}_t{t}_rank"])
        np.testing.assert_allclose(cu.T @ cv, g[f"{name}_t{t}_cu"].T @ g[f"{name}_t{t}_cv"], atol=1e-12)
        dense = b.dense_c(t, alpha, beta)
        if rk == 0:
            assert np.linalg.norm(cu.T @ cv - dense) <= 4 * float(acc.real)


@pytest.mark.parametrize("name,nb,k,T", [("nb500_k25", 500, 25, 3), ("nb1024_k64", 1024, 64, 2)])
def test_oracle_vs_compiled_reference_large(golden_dir, name, nb, k, T):
    g = np.load(os.path.join(golden_dir, "ref_tlr_large.npz"))
    b = make_lr_batch(T, nb, k)
    for t in range(T):
        cu, cv, r, _ = orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0,
                                             beta=1.0, rk=0, maxrk=2 * k, acc=1e-8)
        assert r == int(g[name + "_ranks"][t])
        err = np.linalg.norm(cu.T @ cv - b.dense_c(t, -1.0, 1.0))
        assert abs(err - float(g[name + "_err"][t])) < 1e-11
        assert orc.flop_model(nb, k, k, k, r) == int(g[name + "_flops"][t])


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference():
    rb.set_blas_threads(1)
    b = make_lr_batch(2, 200, 12, base_seed=5)
    for t in range(2):
        cu, cv, r, fl = rb.hcore_gemm(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], 12, alpha=-1.0, beta=1.0,
                                      rk=0, maxrk=24, acc=1e-8)
        cu2, cv2, r2, _ = orc.tlr_gemm_colmajor(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], 12, alpha=-1.0,
                                                beta=1.0, rk=0, maxrk=24, acc=1e-8)
        assert r == r2
        np.testing.assert_allclose(cu[:r], cu2, atol=1e-12)
        np.testing.assert_allclose(cv[:r], cv2, atol=1e-12)
        assert fl == orc.flop_model(200, 12, 12, 12, r)


def test_oracle_error_conventions():
    """Where the reference exit(-1)s the oracle raises (hcore_dgemm.c:743-746, :924-927)."""
    b = make_lr_batch(1, 32, 4)
    args = [x.T for x in (b.au[0], b.av[0], b.bu[0], b.bv[0], b.cu[0][:4], b.cv[0][:4])]
    with pytest.raises(ValueError):
        orc.tlr_gemm(*args, alpha=1.0, beta=1.0, rk=0, maxrk=7, acc=1e-8)
    with pytest.raises(ValueError):
        orc.tlr_gemm(args[0][:, :0], args[1][:, :0], *args[2:], alpha=1.0, beta=1.0, rk=0, maxrk=8, acc=1e-8)


def test_truncation_rule():
    s = np.array([1.0, 1e-3, 1e-9, 1e-12])
    assert orc.truncation_rank(s, 0, 1e-8) == 2
    assert orc.truncation_rank(s, 0, 10.0) == 1          # scan starts at 1: rank >= 1
    assert orc.truncation_rank(s, 0, 1e-20) == 4
    assert orc.truncation_rank(s, 3, 1e-8) == 3
    assert orc.truncation_rank(s, 9, 1e-8) == 4
    assert orc.truncation_rank(s, -1, 1e-8) == min(-1, 4)  # d/s variants: rk != 0 is "fixed" (hcore_dgemm.c:448)
    assert orc.truncation_rank(s, -1, 1e-8, fixed_rank_if_positive=True) == 2   # zgemm: rk > 0 (hcore_zgemm.c:376)


@pytest.mark.parametrize("nb,k", [(64, 8), (150, 10), (300, 20)])
def test_device_model_matches_oracle(nb, k):
    """The numpy model of the DEVICE algorithm (TSQR + QRCP/Jacobi + reflector apply) gives the
    reference's rank and reconstruction error; this validates the math the CUDA kernels implement."""
    b = make_lr_batch(2, nb, k, base_seed=11)
    for t in range(2):
        args = [x.T for x in (b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t][:k], b.cv[t][:k])]
        kw = dict(alpha=-1.0, beta=1.0, rk=0, maxrk=2 * k, acc=1e-8)
        cu, cv, r, sig, sweeps = dm.tlr_gemm_model(*args, **kw)
        cu2, cv2, r2, sig2 = orc.tlr_gemm(*args, **kw)
        D = b.dense_c(t, -1.0, 1.0)
        assert r == r2
        # the pivoted QR drops a trailing block of Frobenius norm < acc / 100: singular values move by less than that
        assert np.abs(sig - sig2).max() < 0.011 * 1e-8 and np.abs(sig - sig2)[sig2 >= 1e-6].max() < 1e-13
        assert np.linalg.norm(cu @ cv.T - D) <= 1.01 * np.linalg.norm(cu2 @ cv2.T - D) + 1e-14
        assert np.abs(cu.T @ cu - np.eye(r)).max() < 1e-13
        assert sweeps <= 8


def test_device_model_rank_deficient_stack():
    """U-stack with exactly repeated columns (C and A share their column space)."""
    b = make_lr_batch(1, 96, 6, base_seed=3)
    au = b.cu[0][:6].T.copy()            # AU == CU
    args = [au, b.av[0].T, b.bu[0].T, b.bv[0].T, b.cu[0][:6].T, b.cv[0][:6].T]
    kw = dict(alpha=-1.0, beta=1.0, rk=0, maxrk=12, acc=1e-8)
    cu, cv, r, sig, _ = dm.tlr_gemm_model(*args, **kw)
    cu2, cv2, r2, _ = orc.tlr_gemm(*args, **kw)
    D = args[4] @ args[5].T - (au @ args[1].T) @ (args[2] @ args[3].T).T
    assert abs(r - r2) <= 1 and r <= 7
    assert np.linalg.norm(cu @ cv.T - D) < 5e-8


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_covariance_tiles_oracle_vs_live_reference():
    """The realistic regime (tiles of one squared-exponential covariance matrix, ragged ranks, rank-deficient
    stack): generator is shard invariant, the oracle agrees with the compiled reference on it."""
    from hcore_b200.tilegen import make_covariance_batch
    rb.set_blas_threads(1)
    nb, maxrk, acc = 160, 96, 1e-8
    b = make_covariance_batch(3, nb, maxrk, acc=acc)
    b2 = make_covariance_batch(1, nb, maxrk, acc=acc, first_tile=2)
    assert np.array_equal(b.cu[2], b2.cu[0]) and np.array_equal(b.av[2], b2.av[0]) and b.brk[2] == b2.brk[0]
    for t in range(3):
        a_, b_, c_ = int(b.ark[t]), int(b.brk[t]), int(b.crk[t])
        cu, cv, r, fl = rb.hcore_gemm(b.au[t, :a_], b.av[t, :a_], b.bu[t, :b_], b.bv[t, :b_], b.cu[t], b.cv[t], c_,
                                      alpha=-1.0, beta=1.0, rk=0, maxrk=maxrk, acc=acc)
        cu2, cv2, r2, _ = orc.tlr_gemm_colmajor(b.au[t, :a_], b.av[t, :a_], b.bu[t, :b_], b.bv[t, :b_], b.cu[t], b.cv[t], c_,
                                                alpha=-1.0, beta=1.0, rk=0, maxrk=maxrk, acc=acc)
        assert r == r2 and r < a_ + c_
        dense = b.dense_c(t, -1.0, 1.0)
        e1, e2 = np.linalg.norm(cu[:r].T @ cv[:r] - dense), np.linalg.norm(cu2.T @ cv2 - dense)
        assert abs(e1 - e2) <= 1e-3 * e1 + 64 * np.finfo(float).eps * np.linalg.norm(dense)


def test_device_model_block_jacobi_and_qrcp():
    """The core-SVD formulation of the device (tests/device_model.py): truncated pivoted QR + block Jacobi on R^T with
    rotations derived from the Gram matrix of a block pair reproduces the singular values of the core to working
    accuracy (graded spectrum: 1 ... 1e-10), keeps the columns orthogonal, and the pair step's J is orthogonal."""
    rng = np.random.default_rng(3)
    n = 40
    u = np.linalg.qr(rng.standard_normal((n, n)))[0]
    v = np.linalg.qr(rng.standard_normal((n, n)))[0]
    sv = 10.0 ** (-10.0 * np.arange(n) / (n - 1))
    t = (u * sv) @ v.T
    r, perm, kq = dm.qrcp(t, tol2=1e-4 * 1e-16)
    assert 0 < kq <= n and sorted(perm.tolist()) == list(range(n))
    q, _ = np.linalg.qr(t[:, perm][:, :kq])
    assert np.linalg.norm(t[:, perm] - q @ r) <= 1e-9 * 1.01 + 1e-13          # dropped block below sqrt(tol2) = 1e-10 (+ slack)
    g, sweeps = dm.jacobi_svd_block4(r.T)
    sig = np.sort(np.sqrt((g * g).sum(0)))[::-1]
    keep = sv[:kq] > 1e-8
    assert np.abs(sig[:kq][keep] - sv[:kq][keep]).max() <= 2e-10               # the truncation perturbs by < 1e-10
    gn = g / np.sqrt((g * g).sum(0))
    assert np.abs(gn.T @ gn - np.eye(kq)).max() < 1e-13 and sweeps <= 10
    x = rng.standard_normal((30, 8)) * 10.0 ** (-rng.integers(0, 8, 8))
    xj, _ = dm.jacobi_pair_step(x, True)
    j = np.linalg.lstsq(x, xj, rcond=None)[0]
    assert np.abs(j.T @ j - np.eye(8)).max() < 1e-10 * np.linalg.cond(x)       # J orthogonal (recovered through x)
    gj = xj.T @ xj
    d = np.sqrt(np.diag(gj))
    c = np.abs(gj) / np.outer(d, d)
    assert np.abs(c[:4, 4:]).max() < 0.5                                       # one pass already reduces the cross cosines


def test_device_model_complex_compact_wy():
    """The complex path's blocked QR and reflector application (tests/device_model.py: zqr_blocked, zapply_blocked; device:
    ztlr_qr_blocked_kernel, ztlr_apply_wy_kernel): 8-reflector blocks through their zlarft factors reproduce the
    unblocked zgeqr2 / zunmqr results (R, tau, Q [y; 0]), including a partial last block and a column that is already
    reduced (tau = 0)."""
    rng = np.random.default_rng(11)
    for m, n, r in ((64, 20, 7), (40, 16, 16), (33, 5, 3)):
        x = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
        if n > 10:
            x[3:, 2] = 0.0                                   # zlarfg: zero tail ...
            x[2, 2] = x[2, 2].real                           # ... and real alpha -> tau = 0 at column 2 unless updates fill it
        xf, tau, tf = dm.zqr_blocked(x)
        # unblocked reference: H_j^H applied column by column to the whole panel
        xr = x.copy()
        taur = np.zeros(n, dtype=complex)
        for j in range(n):
            beta, tj, scale = dm._zlarfg(xr[j, j], float(np.vdot(xr[j + 1:, j], xr[j + 1:, j]).real))
            taur[j] = tj
            if tj != 0.0:
                xr[j, j] = beta
            xr[j + 1:, j] *= scale
            v = np.concatenate(([1.0], xr[j + 1:, j]))
            for c in range(j + 1, n):
                xr[j:, c] -= np.conj(tj) * np.vdot(v, xr[j:, c]) * v
        scale_x = np.linalg.norm(x)
        assert np.abs(tau - taur).max() < 1e-13 and np.abs(xf - xr).max() < 1e-13 * scale_x
        rr = np.triu(xf[:n])
        assert np.allclose(np.abs(rr), np.abs(np.linalg.qr(x)[1]), atol=1e-12 * scale_x)
        # Q [y; 0]: blocked against reflector-by-reflector (H_0 ... H_{n-1} applied last first)
        y = rng.standard_normal((n, r)) + 1j * rng.standard_normal((n, r))
        z = dm.zapply_blocked(xf, tf, y)
        zr = np.zeros((m, r), dtype=complex)
        zr[:n] = y
        for j in range(n - 1, -1, -1):
            v = np.zeros(m, dtype=complex)
            v[j] = 1.0
            v[j + 1:] = xf[j + 1:, j]
            zr -= tau[j] * np.outer(v, v.conj() @ zr)
        assert np.abs(z - zr).max() < 1e-13 * np.linalg.norm(y)
        assert np.abs(z.conj().T @ z - y.conj().T @ y).max() < 1e-12 * np.linalg.norm(y) ** 2   # Q has orthonormal columns
        assert np.abs(z - zr).max() < 1e-12 and np.linalg.norm(xf[:n, :][np.triu_indices(n)]) > 0


def test_device_model_wide_stack_virtual_rows():
    """Stack wider than the TSQR row block (the device's width-256 case: 128-row blocks, up to 256 columns): rows
    h .. n-1 of R are virtual zero rows; R^T R = X^T X, Q has orthonormal columns and Q R = X."""
    rng = np.random.default_rng(3)
    m, n, h, w = 70, 24, 16, 8
    x = rng.random((m, n)) - 0.5
    f = dm.tsqr_factor(x, h=h, w=w)
    R = f["R"]
    assert np.abs(np.tril(R, -1)).max() == 0.0
    assert np.abs(R.T @ R - x.T @ x).max() < 1e-13
    q = dm.tsqr_apply(f, np.eye(n))
    assert np.abs(q.T @ q - np.eye(n)).max() < 1e-13
    assert np.abs(q @ R - x).max() < 1e-13


# ---- the remaining TLR-Cholesky tile kernels + HCORE_dgemm_fast: oracle vs committed compiled-reference outputs ------

def _chol_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_cholstep.npz"))


def test_oracle_trsm_vs_compiled_reference(golden_dir):
    import cholstep_cases as cc
    g = _chol_golden(golden_dir)
    for (side, uplo, tr, dg) in cc.TRSM_VARIANTS:
        a, b = cc.trsm_case(70, 9, side, uplo)
        x, fl = orc.trsm(side, uplo, tr, dg, 0.7, a, b)
        assert np.abs(x - g[f"trsm_{side}_{uplo}_{tr}_{dg}"]).max() < 1e-13
        assert fl == int(g[f"trsm_{side}_{uplo}_{tr}_{dg}_flops"])
    for (m, k, side, uplo, tr) in ((200, 20, 141, 122, 111), (200, 20, 142, 121, 112), (130, 70, 142, 122, 111), (257, 33, 141, 121, 111)):
        a, b = cc.trsm_case(m, k, side, uplo)
        x, _ = orc.trsm(side, uplo, tr, 131, -1.25, a, b)
        assert np.abs(x - g[f"trsmbig_{m}_{k}_{side}_{uplo}_{tr}"]).max() < 1e-12


def test_oracle_potrf_vs_compiled_reference(golden_dir):
    import cholstep_cases as cc
    g = _chol_golden(golden_dir)
    for n in (100, 64, 200):
        for ch in ("L", "U"):
            c, info, fl = orc.potrf(ord(ch), cc.spd_matrix(n))
            assert info == 0 and fl == int(g[f"potrf_{n}_{ch}_flops"])
            assert np.abs(c - g[f"potrf_{n}_{ch}"]).max() < 1e-13
    s_ = cc.spd_matrix(100); s_[50, 50] = -1.0
    c, info, _ = orc.potrf(ord("L"), s_)
    assert info == int(g["potrf_notpd_info"]) == 51
    # the reference forwards the HCORE_enum to LAPACKE's `char uplo` (hcore_dpotrf.c:29-32): the enum value is rejected
    assert orc.potrf(122, cc.spd_matrix(100))[1] == int(g["potrf_enum_info"]) == -2


def test_oracle_gemmbdcd_uncompress_vs_compiled_reference(golden_dir):
    import cholstep_cases as cc
    g = _chol_golden(golden_dir)
    for (m, n, k) in ((96, 96, 7), (130, 80, 12)):
        au, av, bd, cd = cc.bdcd_case(m, n, k)
        for ta in (111, 112):
            for beta in (0.0, 1.0, -0.5):
                out, fl = orc.gemmbdcd(ta, 0.3, au, av, bd, beta, cd)
                assert np.abs(out - g[f"bdcd_{m}_{n}_{k}_{ta}_{beta}"]).max() < 1e-13
                assert fl == int(g[f"bdcd_{m}_{n}_{k}_{ta}_{beta}_flops"])
    for dt, nm in ((np.float64, "d"), (np.complex128, "z")):
        for (m, n, k) in ((96, 80, 7), (130, 200, 33)):
            au, bv = cc.uncompress_case(m, n, k, dt)
            assert np.abs(orc.uncompress(au, bv) - g[f"unc_{nm}_{m}_{n}_{k}"]).max() < 1e-13


def test_oracle_dgemm_fast_vs_compiled_reference(golden_dir):
    """HCORE_dgemm_fast (reference src/compute/hcore_dgemm_fast.c:59-426): same factors as HCORE_dgemm except that
    the accuracy scan starts at index 2 (:314) -- the two crafted tiles get rank 2 where HCORE_dgemm gives 1."""
    import cholstep_cases as cc
    g = _chol_golden(golden_dir)
    for (name, au, av, bu, bv, cu, cv, acc) in cc.fast_cases():
        mk = 2 * max(au.shape[1], cu.shape[1], bu.shape[1])
        cu2, cv2, r, _ = orc.tlr_gemm_fast(au, av, bu, bv, cu, cv, alpha=-1.0, beta=1.0, rk=0, maxrk=mk, acc=acc)
        assert r == int(g[f"fast_{name}_rank"])
        assert np.abs(cu2 @ cv2.T - g[f"fast_{name}_cu"] @ g[f"fast_{name}_cv"].T).max() < 1e-13
        r_gemm = orc.tlr_gemm(au, av, bu, bv, cu, cv, alpha=-1.0, beta=1.0, rk=0, maxrk=mk, acc=acc)[2]
        assert r_gemm == int(g[f"fast_{name}_rank_dgemm"])
    assert int(g["fast_width2_rank"]) == 2 and int(g["fast_width2_rank_dgemm"]) == 1
    assert int(g["fast_width3_rank"]) == 2 and int(g["fast_width3_rank_dgemm"]) == 1


def test_oracle_compress_vs_reference_procedure(golden_dir):
    """oracle.compress against the reference's own compression procedure (testing/testing_dgemm.c:119-236) replayed
    through the LAPACKE of the reference build (tests/golden/make_golden.py compress)."""
    g = np.load(os.path.join(golden_dir, "ref_compress.npz"))
    cases = [(f"c10_{i}", 6, 1e-4) for i in range(3)] + [(f"c{n}", int(g[f"c{n}_par"][0]), float(g[f"c{n}_par"][1])) for n in (64, 96)]
    for name, smr, acc in cases:
        a = g[name + "_a"]
        u, v, r = orc.compress(a, acc, smr)
        assert r == int(g[name + "_rank"])
        assert np.abs(u @ v.T - g[name + "_u"] @ g[name + "_v"].T).max() < 1e-13 * max(1.0, np.abs(a).max())
        assert np.abs(u.T @ u - np.eye(r)).max() < 1e-13
