"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's TLR update path.

A numpy + SciPy-LAPACK restatement of ecrc/hcore's ``HCORE_{d,s,z}gemm`` (low-rank
tile update + recompression), ``HCORE_dgemmcd`` and ``HCORE_dsyrk``.  The reference
delegates every flop to an un-vendored, un-pinned CBLAS/LAPACKE (any found by
``find_package``; reference ``CMakeLists.txt:51,102``); this restatement calls the
same LAPACK routines (``?geqrf``, ``?gesvd``, ``?ormqr``/``zunmqr``) through
``scipy.linalg.lapack`` in the same order, so it reproduces the reference's
results up to BLAS-level rounding.

PARITY PINNING: pinned.  ``tests/test_oracle.py`` checks this file against
(1) the four known-answer programs the reference ships (``testing/testing_*.c``;
their inputs/outputs are committed in ``tests/golden/ref_ctest_*.npz``, produced by
``tests/golden/make_golden.py`` from the unmodified reference built in
``oracle/_ref``), and (2) outputs of the compiled reference on seeded batches at
nb=500/k=25 and nb=1024/k=64 (``tests/golden/ref_*.npz``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Matrices are numpy arrays in *mathematical* orientation
(``x[row, col]``); the ``*_colmajor`` wrappers accept the (k, nb) tile-major layout
used by ``hcore_b200.tilegen``.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import lapack

__all__ = ["qra", "qrb", "core_svd", "newu", "newv", "tlr_gemm", "tlr_gemm_colmajor",
           "gemmcd", "syrk", "flop_model", "truncation_rank",
           "tlr_gemm_fast", "trsm", "potrf", "gemmbdcd", "uncompress", "compress"]

_PFX = {np.dtype(np.float64): "d", np.dtype(np.float32): "s", np.dtype(np.complex128): "z"}


def _lp(name, dt):
    p = _PFX[np.dtype(dt)]
    if p == "z" and name == "ormqr":
        name = "unmqr"
    return getattr(lapack, p + name)


def _geqrf(x):
    qr, tau, _, info = _lp("geqrf", x.dtype)(np.asfortranarray(x))
    assert info == 0
    return qr, tau


def qra(cu, au, alpha, beta):
    """``__qra`` (reference src/compute/hcore_dgemm.c:79-137): QR of ``[beta*CU | alpha*AU]``.

    The reference copies AU behind CU (``dcopy`` :106), scales the two blocks only when
    the scalar differs from one (``dscal`` :109-114) and calls ``dgeqrf`` (:123)."""
    xa = au if alpha == 1 else au * alpha
    xc = cu if beta == 1 else cu * beta
    return _geqrf(np.hstack([xc, xa]))


def qrb(cv, av, bu, bv):
    """``__qrb`` (reference src/compute/hcore_dgemm.c:149-225).

    ``P = AV^T * BV`` (:175), ``G = BU * P^T`` written behind CV (:199-205), ``dgeqrf`` (:216).
    Plain transposes also for complex data (reference src/compute/hcore_zgemm.c:165,189)."""
    p = av.T @ bv
    g = bu @ p.T
    return _geqrf(np.hstack([cv, g]))


def truncation_rank(sigma, rk, acc, fixed_rank_if_positive=False, scan_from=1):
    """Truncation rule of ``__svd`` (reference src/compute/hcore_dgemm.c:444-464).

    ``rk != 0`` -> ``min(rk, n)``; otherwise the first index ``i >= 1`` with
    ``sigma[i] < acc`` (ABSOLUTE threshold, scan starts at 1 so the rank is >= 1), else n.
    ``HCORE_zgemm`` tests ``rk > 0`` instead (reference src/compute/hcore_zgemm.c:376)."""
    n = len(sigma)
    fixed = (rk > 0) if fixed_rank_if_positive else (rk != 0)
    if fixed:
        return min(rk, n)
    for i in range(scan_from, n):          # HCORE_dgemm_fast starts at 2 (reference src/compute/hcore_dgemm_fast.c:314)
        if sigma[i] < acc:
            return i
    return n


def core_svd(qr_u, qr_v, rk, acc, scan_from=1):
    """``__svd`` (reference src/compute/hcore_dgemm.c:296-503).

    ``R_u, R_v`` = upper triangles (:330-336, :358-361); ``T = R_u * R_v^T`` (:384);
    ``gesvd('A','A')`` (:403); truncate (:444-468); scale the first r rows of V^T by sigma
    (:486-489).  Returns ``(U[:, :r], (Sigma V^T)[:r, :], sigma, r)``.
    For complex data LAPACK returns V^H and the reference keeps it as is
    (reference src/compute/hcore_zgemm.c:331-333)."""
    n = qr_u.shape[1]
    ru = np.triu(qr_u[:n, :n])
    rv = np.triu(qr_v[:n, :n])
    t = ru @ rv.T
    u, s, vt, info = _lp("gesvd", t.dtype)(np.asfortranarray(t), compute_uv=1, full_matrices=1)
    assert info == 0
    r = truncation_rank(s, rk, acc, fixed_rank_if_positive=np.iscomplexobj(t), scan_from=scan_from)
    return u[:, :r], vt[:r, :] * s[:r, None], s, r


def newu(qr_u, tau_u, u_r):
    """``__newu`` (reference src/compute/hcore_dgemm.c:508-617): ``CU = Q_u * [U_r; 0]`` via ``dormqr('L','N')`` (:543)."""
    m, n = qr_u.shape
    c = np.zeros((m, u_r.shape[1]), dtype=qr_u.dtype, order="F")
    c[:n] = u_r
    f = _lp("ormqr", qr_u.dtype)
    lwork = max(1, 64 * c.shape[1])
    out, _, info = f("L", "N", qr_u, tau_u, c, lwork)
    assert info == 0
    return out


def newv(qr_v, tau_v, sv_r):
    """``__newv`` (reference src/compute/hcore_dgemm.c:622-696).

    ``[Sigma V^T_r | 0] * Q_v^T`` via ``dormqr('R','T')`` (:655), then transposed into CV (:673).
    Complex: conj, ``zunmqr('R','C')``, transpose, conj back
    (reference src/compute/hcore_zgemm.c:593-632), i.e. ``CV = Q_v * conj(Sigma V^H_r)^T``...
    which this restatement performs literally."""
    m, n = qr_v.shape
    r = sv_r.shape[0]
    c = np.zeros((r, m), dtype=qr_v.dtype, order="F")
    f = _lp("ormqr", qr_v.dtype)
    lwork = max(1, 64 * r)
    if np.iscomplexobj(qr_v):
        c[:, :n] = np.conj(sv_r)
        out, _, info = f("R", "C", qr_v, tau_v, c, lwork)
        assert info == 0
        return np.conj(out.T)
    c[:, :n] = sv_r
    out, _, info = f("R", "T", qr_v, tau_v, c, lwork)
    assert info == 0
    return out.T


def tlr_gemm(au, av, bu, bv, cu, cv, *, alpha, beta, rk, maxrk, acc):
    """``HCORE_dgemm`` (reference src/compute/hcore_dgemm.c:703-1012) on one tile.

    Inputs are ``nb x rank`` matrices (math orientation).  Returns
    ``(cu_new, cv_new, r, sigma)``.  Raises ``ValueError`` where the reference
    ``exit(-1)``s: zero rank (:743-746), ``maxrk < Ark + Crk`` (:924-927),
    new rank above ``maxrk`` (:465-468)."""
    a, b, c = au.shape[1], bu.shape[1], cu.shape[1]
    if a == 0 or b == 0 or c == 0:
        raise ValueError("zero rank")
    if maxrk < a + c:
        raise ValueError("maxrk < Ark + Crk")
    qr_u, tau_u = qra(cu, au, alpha, beta)
    qr_v, tau_v = qrb(cv, av, bu, bv)
    u_r, sv_r, sigma, r = core_svd(qr_u, qr_v, rk, acc)
    if r > maxrk:
        raise ValueError("rank after truncation exceeds maxrk")
    return newu(qr_u, tau_u, u_r), newv(qr_v, tau_v, sv_r), r, sigma


def tlr_gemm_fast(au, av, bu, bv, cu, cv, *, alpha, beta, rk, maxrk, acc):
    """``HCORE_dgemm_fast`` (reference src/compute/hcore_dgemm_fast.c:59-426) on one tile.

    Same stages as ``HCORE_dgemm`` -- stack + ``dgeqrf`` (:150-160, :170-206), core product by ``dtrmm`` on the two R
    factors (:226-229, the same matrix as ``R_u R_v^T``), ``dgesvd`` (:291-300), ``dormqr`` (:355-414) -- with three
    observable differences: the accuracy scan starts at index 2 (:314), there is no ``maxrk`` check on either the stack
    width or the new rank, and no flop counter.  Returns ``(cu_new, cv_new, r, sigma)``."""
    qr_u, tau_u = qra(cu, au, alpha, beta)
    qr_v, tau_v = qrb(cv, av, bu, bv)
    u_r, sv_r, sigma, r = core_svd(qr_u, qr_v, rk, acc, scan_from=2)
    return newu(qr_u, tau_u, u_r), newv(qr_v, tau_v, sv_r), r, sigma


def trsm(side, uplo, trans, diag, alpha, a, b):
    """``HCORE_dtrsm`` (reference src/compute/hcore_dtrsm.c:24-48): ``cblas_dtrsm`` on the factor ``b`` (M x Brank).

    ``side/uplo/trans/diag`` are the HCORE_enum values (141/142, 121/122, 111/112, 131/132).  Returns
    ``(x, flops)`` with the reference's count ``n m^2`` (left) or ``m n^2`` (right) (:42-45, flop_counts 't')."""
    from scipy.linalg import blas
    m, n = b.shape
    x = blas.dtrsm(alpha, np.asfortranarray(a), np.asfortranarray(b), side=int(side == 142), lower=int(uplo == 122),
                   trans_a=int(trans != 111), diag=int(diag == 132))
    return x, (n * m * m if side == 141 else m * n * n)


def potrf(uplo, a):
    """``HCORE_dpotrf`` (reference src/compute/hcore_dpotrf.c:23-35): ``LAPACKE_dpotrf_work`` in place; the other
    triangle is left untouched.  QUIRK: the reference forwards its ``HCORE_enum uplo`` argument unchanged to LAPACKE's
    ``char uplo`` (:29-32), so only the CHARACTER codes 'L' / 'U' (any case) select a triangle; the enum values
    HCORE_Lower = 122 ('z') and HCORE_Upper = 121 ('y') make LAPACK reject the argument: ``LAPACKE_dpotrf_work``
    returns ``info = -2`` (LAPACK's -1, shifted by the LAPACKE wrapper) with the matrix untouched.
    ``uplo`` here is that integer.  Returns ``(a_out, info, flops)``, flops = ``n^3/3 - n^2/2.0 + n/6`` evaluated the
    way ``flop_counts('c', ...)`` does (src/misc/descutil/flop_counts.c:11-14; counted even when info != 0)."""
    n = a.shape[0]
    flops = int(n * n * n // 3 - n * n / 2.0 + n // 6)
    ch = chr(uplo).upper() if 0 < uplo < 128 else "?"
    if ch not in ("L", "U"):
        return np.array(a, copy=True), -2, flops
    c, info = lapack.dpotrf(np.asfortranarray(a), lower=int(ch == "L"), clean=0, overwrite_a=0)
    return c, info, flops


def gemmbdcd(trans_a, alpha, au, av, bd, beta, cd):
    """``HCORE_dgemmbdcd`` (reference src/compute/hcore_dgemmbdcd.c:35-122): ``CD <- alpha AU (AV^T BD) + beta CD``
    (:78-90 / :99-111; the two ``transA`` branches differ only in which of M, LDA, LDB they use as sizes and agree
    when LDA = LDB = M).  Returns ``(cd_out, flops)`` with ``flops->update`` as the reference counts it (:91-92)."""
    k, n, m = au.shape[1], bd.shape[1], au.shape[0]
    w = av.T @ bd
    out = alpha * (au @ w) + (beta * cd if beta != 0 else 0.0)
    return out, 2 * k * n * bd.shape[0] + 2 * m * n * k


def uncompress(au, bv):
    """``HCORE_duncompress`` / ``HCORE_zuncompress`` (reference src/compute/hcore_duncompress.c:33-85,
    hcore_zuncompress.c:35-95): ``CD = AU BV^T`` with a PLAIN transpose also for complex data; the reference passes
    one / zero to ``?gemm`` whatever ``alpha`` / ``beta`` the caller gives (:72-78) and asserts equal ranks (:69)."""
    assert au.shape[1] == bv.shape[1], "ranks must be the same"
    return au @ bv.T


def compress(a, acc, storage_maxrank):
    """Dense tile -> ``(U, V, rank)`` the way the reference's tests do it (reference testing/testing_dgemm.c:119-236; the
    library itself only declares ``HCORE_zgytlr``, src/include/hcore_z.h:63-75, and never defines it): ``dgesvd('A','A')``
    (:172), rows of ``V^T`` scaled by sigma (:176-179), ``rank = #{sigma_i >= acc}`` counted from i = 0 (:201), capped at
    ``storage_maxrank`` (:202-205), ``U`` = first rank left vectors (:212), ``V = (Sigma V^T)^T`` (:217).
    Returns math-orientation ``U`` (m x rank), ``V`` (n x rank) with ``a ~ U @ V.T``."""
    u, s, vt, info = lapack.dgesvd(np.asfortranarray(a), compute_uv=1, full_matrices=1)
    assert info == 0
    rank = 0
    while rank < min(a.shape) and s[rank] >= acc:
        rank += 1
    rank = min(rank, storage_maxrank)
    return u[:, :rank].copy(), (vt[:rank] * s[:rank, None]).T.copy(), rank


def tlr_gemm_colmajor(au, av, bu, bv, cu, cv, crk, **kw):
    """Same, for (k, nb) tile-major arrays (``hcore_b200.tilegen`` layout). Returns (r, nb)-shaped factors."""
    cu_new, cv_new, r, sigma = tlr_gemm(au.T, av.T, bu.T, bv.T, cu[:crk].T, cv[:crk].T, **kw)
    return np.ascontiguousarray(cu_new.T), np.ascontiguousarray(cv_new.T), r, sigma


def gemmcd(au, av, bu, bv, cd):
    """``HCORE_dgemmcd`` (reference src/compute/hcore_dgemmcd.c:36-107).

    ``CD <- CD - (Au Av^T)(Bu Bv^T)^T`` with alpha/beta hard-wired to -1/+1; the GEMM
    order depends on ``Arank > Brank`` (:72)."""
    w = av.T @ bv
    if au.shape[1] > bu.shape[1]:
        return cd - (au @ w) @ bu.T
    return cd - au @ (w @ bu.T)


def syrk(au, av, cd, alpha):
    """``HCORE_dsyrk`` (reference src/compute/hcore_dsyrk.c:47-189).

    ``CD <- CD + alpha * AU (AV^T AV) AU^T`` (beta ignored, :165-171), then the strict
    upper triangle is set to zero (:174-181)."""
    w1 = av.T @ av
    w2 = au @ w1
    out = cd + alpha * (w2 @ au.T)
    return np.tril(out)


def flop_model(m, a, b, c, r, include_core=False):
    """The reference's own flop accounting for one ``HCORE_dgemm`` call
    (``flops->update``; reference src/misc/descutil/flop_counts.c:3-70 and the call sites
    src/compute/hcore_dgemm.c:126,179,208,224,434,493,595,666).  The ``T = R_u R_v^T``
    product (:384-386) is not counted by the reference; ``include_core`` adds its 2n^3."""
    n = a + c

    def geqrf(mm, nn):
        return 2 * mm * nn * nn - int(np.float32(2 * nn * nn * nn) / np.float32(3.0)) + 2 * mm * nn + int(np.float32(17 * nn) / np.float32(3.0))

    f = geqrf(m, n)                       # __qra
    f += 2 * a * b * m + 2 * m * a * b    # __qrb GEMMs
    f += geqrf(m, n)                      # __qrb QR
    f += 22 * n ** 3                      # gesvd model
    f += r * n                            # sigma scaling
    f += 4 * r * m * n - 2 * r * n * n + 3 * r * n                                    # ormqr left  (m,n,k)=(m,r,n)
    kk, nn_, mm_ = n, m, r                                                              # ormqr right (m,n,k)=(r,m,n)
    f += 4 * nn_ * mm_ * kk - 2 * mm_ * kk * kk + 2 * mm_ * kk + nn_ * kk - kk * kk // 2 + kk // 2
    if include_core:
        f += 2 * n ** 3
    return f
