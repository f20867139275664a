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
           "gemmcd", "syrk", "flop_model", "truncation_rank"]

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


def truncation_rank(sigma, rk, acc, fixed_rank_if_positive=False):
    """Truncation rule of ``__svd`` (reference src/compute/hcore_dgemm.c:444-464).

    ``rk != 0`` -> ``min(rk, n)``; otherwise the first index ``i >= 1`` with
    ``sigma[i] < acc`` (ABSOLUTE threshold, scan starts at 1 so the rank is >= 1), else n.
    ``HCORE_zgemm`` tests ``rk > 0`` instead (reference src/compute/hcore_zgemm.c:376)."""
    n = len(sigma)
    fixed = (rk > 0) if fixed_rank_if_positive else (rk != 0)
    if fixed:
        return min(rk, n)
    for i in range(1, n):
        if sigma[i] < acc:
            return i
    return n


def core_svd(qr_u, qr_v, rk, acc):
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
    r = truncation_rank(s, rk, acc, fixed_rank_if_positive=np.iscomplexobj(t))
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
