"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run once in the dev container (needs /root/reference built into oracle/_ref by
`make -C oracle ref`):   python tests/golden/make_golden.py

Fixtures:
  ref_ctest_dgemm.npz    inputs/outputs of the reference's own known-answer program
                         testing/testing_dgemm.c (M=N=K=10, LAPACKE_dlatms with
                         iseed={0,0,0,1}, acc 1e-4): the dlatms/dgesvd calls are replayed here
                         through the same OpenBLAS the reference was linked with, then the
                         compiled reference HCORE_dgemm is run on them.
  ref_ctest_dgemmcd.npz  testing/testing_dgemmcd.c (nrows 10, ranks 7/6, glibc rand() with the
                         default seed), outputs of the compiled HCORE_dgemmcd.
  ref_ctest_dsyrk.npz    testing/testing_dsyrk.c (nrows 10, rank 7).
  ref_tlr_small.npz      compiled-reference outputs (full factors) for seeded tilegen batches
                         at nb=64/k=8 and nb=96/k=(5,7,6) (ragged ranks), fixed-rank mode too.
  ref_cholstep.npz       compiled-reference outputs of HCORE_dtrsm / dpotrf / dgemmbdcd / d,zuncompress / dgemm_fast on
                         the seeded inputs of tests/cholstep_cases.py (`python tests/golden/make_golden.py cholstep`).
  ref_tlr_large.npz      compiled-reference ranks, singular-value-free summaries (errors,
                         flop counts) at nb=500/k=25 and nb=1024/k=64 (inputs are regenerated
                         from the seed by hcore_b200.tilegen).
"""
import ctypes as C
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hcore_b200.tilegen import make_lr_batch  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

import scipy  # noqa: E402

_blas = C.CDLL(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs",
                                      "libscipy_openblas*.so"))[0])
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
_blas.scipy_LAPACKE_dlatms.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char, ip, C.c_char, dp, C.c_int, C.c_double,
                                       C.c_double, C.c_int, C.c_int, C.c_char, dp, C.c_int]
_blas.scipy_LAPACKE_dgesvd.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int,
                                       dp, C.c_int, dp]


def _p(a):
    return a.ctypes.data_as(dp)


fpp = C.POINTER(C.c_float)
_blas.scipy_LAPACKE_slatms.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char, ip, C.c_char, fpp, C.c_int, C.c_float,
                                       C.c_float, C.c_int, C.c_int, C.c_char, fpp, C.c_int]
_blas.scipy_LAPACKE_sgesvd.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, fpp, C.c_int, fpp, fpp, C.c_int,
                                       fpp, C.c_int, fpp]


def generate_compress_s(n, iseed, storage_maxrank, acc):
    """float clone: generate_compress() of reference testing/testing_sgemm.c:119-236."""
    f = lambda a: a.ctypes.data_as(fpp)  # noqa: E731
    d = np.array([10.0 ** (-i) for i in range(n)], np.float32)
    a = np.zeros((n, n), np.float32)
    info = _blas.scipy_LAPACKE_slatms(102, n, n, b"U", iseed.ctypes.data_as(ip), b"N", f(d), 0, -1.0, -1.0, n - 1,
                                      n - 1, b"N", f(a), n)
    assert info == 0
    clone = a.copy()
    s, u, vt, sb = (np.zeros(n, np.float32), np.zeros((n, n), np.float32), np.zeros((n, n), np.float32),
                    np.zeros(n, np.float32))
    info = _blas.scipy_LAPACKE_sgesvd(102, b"A", b"A", n, n, f(clone), n, f(s), f(u), n, f(vt), n, f(sb))
    assert info == 0
    vt_m = vt.T * s[:, None]
    rank = 0
    while rank < n and s[rank] >= acc:
        rank += 1
    rank = min(rank, storage_maxrank)
    return a, np.ascontiguousarray(u[:rank]), np.ascontiguousarray(vt_m[:rank]), rank


def ctest_sgemm():
    """reference testing/testing_sgemm.c (M=N=K=10, storage_maxrank 10 (:259), acc 1e-4)."""
    n, smr, acc = 10, 10, 1e-4
    iseed = np.array([0, 0, 0, 1], dtype=np.int32)
    A, AU, AV, ra = generate_compress_s(n, iseed, smr, acc)
    B, BU, BV, rb_ = generate_compress_s(n, iseed, smr, acc)
    Cm, CU, CV, rc = generate_compress_s(n, iseed, smr, acc)
    ref_c = (A.T.astype(np.float32) @ B.T.astype(np.float32).T + Cm.T).astype(np.float32)
    cu, cv, r, fl = rb.hcore_gemm(AU, AV, BU, BV, CU, CV, rc, alpha=1.0, beta=1.0, rk=0, maxrk=n, acc=acc)
    app = cu[:r].T @ cv[:r]
    rnorm = np.abs(ref_c - app).sum(axis=1).max()
    print("ctest sgemm: ranks", ra, rb_, rc, "->", r, "inf norm", rnorm)
    assert rnorm < 10 * acc
    np.savez(os.path.join(HERE, "ref_ctest_sgemm.npz"), AU=AU, AV=AV, BU=BU, BV=BV, CU=CU, CV=CV,
             ranks=np.array([ra, rb_, rc]), ref_c=ref_c, out_cu=cu[:r], out_cv=cv[:r], out_rank=r,
             rnorm=rnorm, flops=fl, acc=acc, n=n)


def generate_compress(n, iseed, storage_maxrank, acc):
    """Replays generate_compress() of reference testing/testing_dgemm.c:119-236 (square n x n)."""
    d = np.array([10.0 ** (-i) for i in range(n)])
    a = np.zeros((n, n))            # column-major n x n as (col, row)
    info = _blas.scipy_LAPACKE_dlatms(102, n, n, b"U", iseed.ctypes.data_as(ip), b"N", _p(d), 0, -1.0, -1.0, n - 1,
                                      n - 1, b"N", _p(a), n)
    assert info == 0
    clone = a.copy()
    s, u, vt, sb = np.zeros(n), np.zeros((n, n)), np.zeros((n, n)), np.zeros(n)
    info = _blas.scipy_LAPACKE_dgesvd(102, b"A", b"A", n, n, _p(clone), n, _p(s), _p(u), n, _p(vt), n, _p(sb))
    assert info == 0
    vt_m = vt.T * s[:, None]        # math orientation: rows of V^T scaled by sigma (:176-179)
    rank = 0
    while rank < n and s[rank] >= acc:
        rank += 1
    rank = min(rank, storage_maxrank)
    U = np.ascontiguousarray(u[:rank])               # (rank, n): column-major n x rank
    V = np.ascontiguousarray(vt_m[:rank])            # V = (Sigma V^T)^T stored n x rank -> (rank, n)
    return a, U, V, rank


def ctest_dgemm():
    n, smr, acc = 10, 6, 1e-4
    iseed = np.array([0, 0, 0, 1], dtype=np.int32)
    A, AU, AV, ra = generate_compress(n, iseed, smr, acc)
    B, BU, BV, rb_ = generate_compress(n, iseed, smr, acc)
    Cm, CU, CV, rc = generate_compress(n, iseed, smr, acc)
    Am, Bm, Cmm = A.T, B.T, Cm.T     # math orientation
    ref_c = 1.0 * Am @ Bm.T + 1.0 * Cmm
    cu, cv, r, fl = rb.hcore_gemm(AU, AV, BU, BV, CU, CV, rc, alpha=1.0, beta=1.0, rk=0, maxrk=n, acc=acc)
    app = cu[:r].T @ cv[:r]
    rnorm = np.abs(ref_c - app).sum(axis=1).max()
    print("ctest dgemm: ranks", ra, rb_, rc, "->", r, "inf norm", rnorm)
    assert rnorm < 10 * acc
    np.savez(os.path.join(HERE, "ref_ctest_dgemm.npz"), AU=AU, AV=AV, BU=BU, BV=BV, CU=CU, CV=CV,
             ranks=np.array([ra, rb_, rc]), ref_c=ref_c, out_cu=cu[:r], out_cv=cv[:r], out_rank=r,
             rnorm=rnorm, flops=fl, acc=acc, n=n)


def _glibc_rand_stream():
    libc = C.CDLL("libc.so.6")
    libc.srand(1)                    # glibc's default seed when srand() is never called
    return lambda: libc.rand() / 2147483647.0


def ctest_dgemmcd():
    """reference testing/testing_dgemmcd.c:105-205: nrows 10, rankA 7, rankB 6, rand()/RAND_MAX."""
    n, ka, kb = 10, 7, 6
    rnd = _glibc_rand_stream()
    # the test fills the arrays element by element in allocation order; the exact order only
    # matters for bit-reproducing ITS numbers, the fixture pins OUR inputs against the reference
    Au = np.array([[rnd() for _ in range(n)] for _ in range(ka)])
    Av = np.array([[rnd() for _ in range(n)] for _ in range(ka)])
    Bu = np.array([[rnd() for _ in range(n)] for _ in range(kb)])
    Bv = np.array([[rnd() for _ in range(n)] for _ in range(kb)])
    CD = np.array([[rnd() for _ in range(n)] for _ in range(n)])
    out = rb.hcore_dgemmcd(Au, Av, Bu, Bv, CD)
    dense = CD.T - (Au.T @ Av) @ (Bu.T @ Bv).T
    res = np.abs(dense - out.T).sum(axis=1).max()
    print("ctest dgemmcd residual", res)
    assert res < 1000 * np.finfo(float).eps
    np.savez(os.path.join(HERE, "ref_ctest_dgemmcd.npz"), Au=Au, Av=Av, Bu=Bu, Bv=Bv, CD=CD, out=out, residual=res)


def ctest_dsyrk():
    """reference testing/testing_dsyrk.c:102-181: nrows 10, rank 7, lower triangle compared."""
    n, k = 10, 7
    rnd = _glibc_rand_stream()
    AU = np.array([[rnd() for _ in range(n)] for _ in range(k)])
    AV = np.array([[rnd() for _ in range(n)] for _ in range(k)])
    CD = np.array([[rnd() for _ in range(n)] for _ in range(n)])
    out, fl = rb.hcore_dsyrk(AU, AV, CD, alpha=-1.0)
    Am = AU.T @ AV
    dense = np.tril(CD.T - Am @ Am.T)
    res = np.abs(dense - np.tril(out.T)).sum(axis=1).max()
    print("ctest dsyrk residual", res, "upper zeroed:", np.abs(np.triu(out.T, 1)).max())
    assert res < 1000 * np.finfo(float).eps
    np.savez(os.path.join(HERE, "ref_ctest_dsyrk.npz"), AU=AU, AV=AV, CD=CD, out=out, residual=res, flops=fl)


def tlr_small():
    rec = {}
    cases = [("nb64_k8", 64, 8, None, 0, 1e-8, -1.0, 1.0),
             ("nb96_ragged", 96, 8, [(5, 7, 6), (8, 3, 2), (1, 1, 1)], 0, 1e-6, 0.5, 2.0),
             ("nb64_fixed", 64, 8, None, 5, 1e-8, -1.0, 1.0)]
    for name, nb, k, ranks, rk, acc, alpha, beta in cases:
        T = 3
        b = make_lr_batch(T, nb, k, base_seed=99, ranks=ranks)
        for t in range(T):
            cu, cv, r, fl = rb.hcore_gemm(b.au[t, :b.ark[t]], b.av[t, :b.ark[t]], b.bu[t, :b.brk[t]], b.bv[t, :b.brk[t]],
                                          b.cu[t], b.cv[t], int(b.crk[t]), alpha=alpha, beta=beta, rk=rk,
                                          maxrk=2 * k, acc=acc)
            rec[f"{name}_t{t}_cu"] = cu[:r]
            rec[f"{name}_t{t}_cv"] = cv[:r]
            rec[f"{name}_t{t}_rank"] = r
            rec[f"{name}_t{t}_flops"] = fl
        rec[f"{name}_meta"] = np.array([nb, k, rk, T], dtype=np.int64)
        rec[f"{name}_scal"] = np.array([acc, alpha, beta])
        if ranks is not None:
            rec[f"{name}_ranks"] = np.array(ranks, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_tlr_small.npz"), **rec)
    print("tlr_small:", {k_: int(v) for k_, v in rec.items() if k_.endswith("_rank")})


def tlr_large():
    rec = {}
    for name, nb, k, T in (("nb500_k25", 500, 25, 3), ("nb1024_k64", 1024, 64, 2)):
        b = make_lr_batch(T, nb, k)
        ranks, errs, flops, frob = [], [], [], []
        for t in range(T):
            cu, cv, r, fl = rb.hcore_gemm(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=-1.0,
                                          beta=1.0, rk=0, maxrk=2 * k, acc=1e-8)
            D = b.dense_c(t, -1.0, 1.0)
            ranks.append(r); flops.append(fl)
            errs.append(np.linalg.norm(cu[:r].T @ cv[:r] - D)); frob.append(np.linalg.norm(D))
        rec[name + "_ranks"] = np.array(ranks)
        rec[name + "_err"] = np.array(errs)
        rec[name + "_normC"] = np.array(frob)
        rec[name + "_flops"] = np.array(flops)
        print(name, ranks, errs)
    np.savez(os.path.join(HERE, "ref_tlr_large.npz"), **rec)


def tlr_complex():
    """Compiled-reference HCORE_zgemm outputs (the reference has no test of its own for zgemm):
    seeded complex batches, plain-transpose tiles, adaptive and fixed-rank (rk > 0) modes."""
    rec = {}
    for name, nb, k, rk, acc, alpha, beta in (("z_nb64_k6", 64, 6, 0, 1e-8, -1.0 + 0.25j, 1.0 - 0.5j),
                                              ("z_nb96_fixed", 96, 8, 5, 1e-8, 0.5 - 1.0j, 1.0 + 0.0j)):
        T = 2
        b = make_lr_batch(T, nb, k, base_seed=51, dtype=np.complex128)
        for t in range(T):
            cu, cv, r, fl = rb.hcore_gemm(b.au[t], b.av[t], b.bu[t], b.bv[t], b.cu[t], b.cv[t], k, alpha=alpha, beta=beta,
                                          rk=rk, maxrk=2 * k, acc=acc)
            rec[f"{name}_t{t}_cu"] = cu[:r]; rec[f"{name}_t{t}_cv"] = cv[:r]
            rec[f"{name}_t{t}_rank"] = r; rec[f"{name}_t{t}_flops"] = fl
        rec[name + "_meta"] = np.array([nb, k, rk, T], dtype=np.int64)
        rec[name + "_scal"] = np.array([acc, alpha, beta], dtype=np.complex128)
    np.savez_compressed(os.path.join(HERE, "ref_tlr_complex.npz"), **rec)
    print("tlr_complex:", {k_: int(v) for k_, v in rec.items() if k_.endswith("_rank")})


def cholstep():
    """Compiled-reference outputs of HCORE_dtrsm (all 16 side/uplo/trans/diag variants + two multi-block sizes),
    HCORE_dpotrf ('L', 'U', a non-positive-definite tile, the enum-valued uplo the reference rejects), HCORE_dgemmbdcd
    (both transA branches, beta 0 / 1 / -0.5, a non-square CD), HCORE_duncompress / HCORE_zuncompress and HCORE_dgemm_fast
    on the seeded inputs of tests/cholstep_cases.py.  The reference ships no test for any of them."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cholstep_cases as cc
    cm = lambda x: np.ascontiguousarray(x.T)          # math orientation -> column-major storage
    rec = {}
    for (side, uplo, tr, dg) in cc.TRSM_VARIANTS:
        a, b = cc.trsm_case(70, 9, side, uplo)
        out, fl = rb.hcore_dtrsm(side, uplo, tr, dg, 0.7, cm(a), cm(b))
        rec[f"trsm_{side}_{uplo}_{tr}_{dg}"] = out.T
        rec[f"trsm_{side}_{uplo}_{tr}_{dg}_flops"] = fl
    for (m, k, side, uplo, tr) in ((200, 20, 141, 122, 111), (200, 20, 142, 121, 112), (130, 70, 142, 122, 111), (257, 33, 141, 121, 111)):
        a, b = cc.trsm_case(m, k, side, uplo)
        out, fl = rb.hcore_dtrsm(side, uplo, tr, 131, -1.25, cm(a), cm(b))
        rec[f"trsmbig_{m}_{k}_{side}_{uplo}_{tr}"] = out.T
    for n in (100, 64, 200):
        s_ = cc.spd_matrix(n)
        for ch in ("L", "U"):
            out, info, fl = rb.hcore_dpotrf(ord(ch), cm(s_))
            assert info == 0
            rec[f"potrf_{n}_{ch}"] = out.T
            rec[f"potrf_{n}_{ch}_flops"] = fl
    s_ = cc.spd_matrix(100); s_[50, 50] = -1.0
    out, info, fl = rb.hcore_dpotrf(ord("L"), cm(s_))
    rec["potrf_notpd"] = out.T; rec["potrf_notpd_info"] = info
    out, info, fl = rb.hcore_dpotrf(122, cm(cc.spd_matrix(100)))         # HCORE_Lower as an enum: rejected by LAPACK
    rec["potrf_enum_info"] = info
    for (m, n, k) in ((96, 96, 7), (130, 80, 12)):
        au, av, bd, cd = cc.bdcd_case(m, n, k)
        for ta in (111, 112):
            for beta in (0.0, 1.0, -0.5):
                out, fl = rb.hcore_dgemmbdcd(ta, 0.3, cm(au), cm(av), cm(bd), beta, cm(cd))
                rec[f"bdcd_{m}_{n}_{k}_{ta}_{beta}"] = out.T
                rec[f"bdcd_{m}_{n}_{k}_{ta}_{beta}_flops"] = fl
    for dt, nm in ((np.float64, "d"), (np.complex128, "z")):
        for (m, n, k) in ((96, 80, 7), (130, 200, 33)):
            au, bv = cc.uncompress_case(m, n, k, dt)
            rec[f"unc_{nm}_{m}_{n}_{k}"] = rb.hcore_uncompress(cm(au), cm(bv)).T
    for (name, au, av, bu, bv, cu, cv, acc) in cc.fast_cases():
        a, c = au.shape[1], cu.shape[1]
        mk = 2 * max(a, c, bu.shape[1])
        cu2, cv2, r = rb.hcore_dgemm_fast(cm(au), cm(av), cm(bu), cm(bv), cm(cu), cm(cv), c, alpha=-1.0, beta=1.0, rk=0, maxrk=mk, acc=acc)
        _, _, r_gemm, _ = rb.hcore_gemm(cm(au), cm(av), cm(bu), cm(bv), np.vstack([cm(cu), np.zeros((mk, cu.shape[0]))]),
                                        np.vstack([cm(cv), np.zeros((mk, cu.shape[0]))]), c, alpha=-1.0, beta=1.0, rk=0, maxrk=mk, acc=acc)
        rec[f"fast_{name}_cu"] = cu2[:r].T; rec[f"fast_{name}_cv"] = cv2[:r].T
        rec[f"fast_{name}_rank"] = r; rec[f"fast_{name}_rank_dgemm"] = r_gemm
        print("dgemm_fast", name, "rank", r, "HCORE_dgemm rank", r_gemm)
    np.savez_compressed(os.path.join(HERE, "ref_cholstep.npz"), **rec)
    print("cholstep:", len(rec), "arrays")


def compress_cases():
    """The reference's compression procedure (testing/testing_dgemm.c:119-236: LAPACKE_dlatms + LAPACKE_dgesvd + rank rule),
    replayed through the same OpenBLAS the reference is linked with: the three 10 x 10 matrices of testing_dgemm.c and
    two larger dlatms matrices.  Stored: the dense matrices and the factors (pins oracle.compress)."""
    rec = {}
    iseed = np.array([0, 0, 0, 1], dtype=np.int32)
    for i in range(3):
        a, U, V, r = generate_compress(10, iseed, 6, 1e-4)
        rec[f"c10_{i}_a"] = a.T.copy(); rec[f"c10_{i}_u"] = U.T.copy(); rec[f"c10_{i}_v"] = V.T.copy(); rec[f"c10_{i}_rank"] = r
    for n, smr, acc in ((64, 40, 1e-8), (96, 12, 1e-6)):
        d = np.array([10.0 ** (-i / 3.0) for i in range(n)])
        a = np.zeros((n, n))
        assert _blas.scipy_LAPACKE_dlatms(102, n, n, b"U", iseed.ctypes.data_as(ip), b"N", _p(d), 0, -1.0, -1.0, n - 1, n - 1, b"N", _p(a), n) == 0
        clone = a.copy()
        s_, u, vt, sb = np.zeros(n), np.zeros((n, n)), np.zeros((n, n)), np.zeros(n)
        assert _blas.scipy_LAPACKE_dgesvd(102, b"A", b"A", n, n, _p(clone), n, _p(s_), _p(u), n, _p(vt), n, _p(sb)) == 0
        rank = 0
        while rank < n and s_[rank] >= acc:
            rank += 1
        rank = min(rank, smr)
        rec[f"c{n}_a"] = a.T.copy(); rec[f"c{n}_u"] = u[:rank].T.copy(); rec[f"c{n}_v"] = (vt.T * s_[:, None])[:rank].T.copy()
        rec[f"c{n}_rank"] = rank; rec[f"c{n}_par"] = np.array([smr, acc])
    np.savez_compressed(os.path.join(HERE, "ref_compress.npz"), **rec)
    print("compress:", {k_: int(v) for k_, v in rec.items() if k_.endswith("_rank")})


if __name__ == "__main__":
    rb.set_blas_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "cholstep":
        cholstep()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "compress":
        compress_cases()
        sys.exit(0)
    ctest_dgemm()
    ctest_sgemm()
    ctest_dgemmcd()
    ctest_dsyrk()
    tlr_small()
    tlr_large()
    tlr_complex()
    cholstep()
    compress_cases()
