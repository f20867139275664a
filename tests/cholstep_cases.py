"""Seeded inputs of the parity cases for the remaining TLR-Cholesky tile kernels (HCORE_dtrsm, HCORE_dpotrf,
HCORE_dgemmbdcd, HCORE_d/zuncompress) and HCORE_dgemm_fast.  Shared by tests/golden/make_golden.py (which runs the
compiled reference on them), tests/test_oracle.py (oracle vs those outputs) and tests/test_gpu_cholstep.py (CUDA path vs
both).  Matrices are in mathematical orientation (x[row, col])."""
import itertools

import numpy as np

from hcore_b200.tilegen import make_lr_batch

TRSM_VARIANTS = list(itertools.product((141, 142), (121, 122), (111, 112), (131, 132)))   # side, uplo, trans, diag


def tri_matrix(rng, n, uplo):
    """Well-conditioned triangular matrix; the other triangle holds garbage that must never be read."""
    a = (rng.random((n, n)) - 0.5) / np.sqrt(n)
    a[np.arange(n), np.arange(n)] = 1.0 + rng.random(n)
    junk = 1e6 * (rng.random((n, n)) - 0.5)
    return np.tril(a) + np.triu(junk, 1) if uplo == 122 else np.triu(a) + np.tril(junk, -1)


def trsm_case(m, k, side, uplo, seed=5):
    rng = np.random.default_rng(seed + 7 * m + k)
    a = tri_matrix(rng, m if side == 141 else k, uplo)
    b = rng.random((m, k)) - 0.5
    return a, b


def spd_matrix(n, seed=3):
    rng = np.random.default_rng(seed + n)
    x = rng.random((n, n)) - 0.5
    return x @ x.T / n + np.eye(n)


def bdcd_case(m, n, k, seed=11):
    b = make_lr_batch(1, m, k, base_seed=seed)
    rng = np.random.default_rng(seed + m + n)
    return b.au[0].T.copy(), b.av[0].T.copy(), rng.random((m, n)) - 0.5, rng.random((m, n)) - 0.5


def uncompress_case(m, n, k, dtype, seed=13):
    rng = np.random.default_rng(seed + m + n + k)
    f = (lambda s: rng.random(s) - 0.5 + 1j * (rng.random(s) - 0.5)) if dtype == np.complex128 else (lambda s: rng.random(s) - 0.5)
    return f((m, k)).astype(dtype), f((n, k)).astype(dtype)


def fast_cases():
    """(name, au, av, bu, bv, cu, cv, acc) for HCORE_dgemm_fast: two generic tiles and two where the accuracy scan
    starting at index 2 (reference src/compute/hcore_dgemm_fast.c:314) gives a larger rank than HCORE_dgemm's scan from 1."""
    out = []
    b = make_lr_batch(2, 96, 7, base_seed=61)
    for t in range(2):
        out.append((f"generic{t}", b.au[t].T.copy(), b.av[t].T.copy(), b.bu[t].T.copy(), b.bv[t].T.copy(),
                    b.cu[t][:7].T.copy(), b.cv[t][:7].T.copy(), 1e-8))
    rng = np.random.default_rng(71)
    q = np.linalg.qr(rng.random((64, 4)) - 0.5)[0]
    w = np.linalg.qr(rng.random((64, 4)) - 0.5)[0]
    # stack of width 2 whose second singular value is ~0: HCORE_dgemm -> rank 1, HCORE_dgemm_fast -> rank 2
    out.append(("width2", q[:, :1].copy(), w[:, :1].copy(), w[:, 1:2].copy(), w[:, :1].copy(), q[:, :1].copy(), w[:, 2:3].copy(), 1e-8))
    # stack of width 3, numerical rank 1: HCORE_dgemm -> 1, HCORE_dgemm_fast -> 2
    cu = q[:, :2].copy(); cv = w[:, :2] * np.array([1.0, 1e-12])
    out.append(("width3", q[:, :1].copy(), w[:, 2:3].copy(), w[:, 3:4].copy(), w[:, 2:3].copy(), cu, cv, 1e-8))
    return out
