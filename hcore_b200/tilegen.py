"""Synthetic tile-low-rank (TLR) batches for tests and benchmarks.

Follows the generator SURVEY.md section 8(d) defines for the hot path (the
reference itself ships no generator beyond its 10x10 ``dlatms`` test, reference
``testing/testing_dgemm.c:119-236``):

* ``U`` factors have orthonormal columns, ``V`` factors carry the singular
  values (HiCMA convention, reference ``testing/testing_dgemm.c:176-179,206,211``):
  ``V = orth(rand) * diag(sigma)`` with ``sigma_i = 10**(-decades * i / k)``.
* every tile ``t`` is seeded with ``base_seed + t`` so a tile's content does not
  depend on how the batch is sharded over GPUs.

All arrays are column-major per tile: ``a[t]`` has shape ``(k, nb)`` in numpy
(C order) which is exactly an ``nb x k`` column-major matrix with ``ld = nb``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["orth_factor", "make_lr_batch", "make_dense_batch", "make_covariance_batch", "TlrBatch"]


def orth_factor(rng: np.random.Generator, nb: int, k: int, dtype=np.float64) -> np.ndarray:
    """``nb x k`` matrix with orthonormal columns, returned column-major as (k, nb)."""
    if np.issubdtype(dtype, np.complexfloating):
        x = rng.random((nb, k)) - 0.5 + 1j * (rng.random((nb, k)) - 0.5)
    else:
        x = rng.random((nb, k)) - 0.5
    q, _ = np.linalg.qr(x)
    return np.ascontiguousarray(q.T).astype(dtype)


class TlrBatch:
    """A batch of independent update problems ``C <- alpha*A*B^T + beta*C``.

    Attributes (numpy, tile-major, each tile column-major with ld = nb):
      au, av : (T, ka, nb)      bu, bv : (T, kb, nb)
      cu, cv : (T, maxrk, nb)   -- first ``crk[t]`` columns valid, rest zero
      ark, brk, crk : (T,) int32
    """

    def __init__(self, nb, maxrk, au, av, bu, bv, cu, cv, ark, brk, crk):
        self.nb, self.maxrk = nb, maxrk
        self.au, self.av, self.bu, self.bv, self.cu, self.cv = au, av, bu, bv, cu, cv
        self.ark, self.brk, self.crk = ark, brk, crk

    @property
    def ntiles(self):
        return self.au.shape[0]

    def dense_c(self, t, alpha, beta):
        """Dense reference ``alpha*A*B^T + beta*C`` for tile ``t`` (nb x nb, numpy row/col = matrix row/col).

        For complex data the tile is ``U * V^T`` with a *plain* transpose
        (reference ``src/compute/hcore_zgemm.c:165,189,309``)."""
        a, b, c = int(self.ark[t]), int(self.brk[t]), int(self.crk[t])
        AU, AV = self.au[t, :a].T, self.av[t, :a].T
        BU, BV = self.bu[t, :b].T, self.bv[t, :b].T
        CU, CV = self.cu[t, :c].T, self.cv[t, :c].T
        A = AU @ AV.T
        B = BU @ BV.T
        return alpha * (A @ B.T) + beta * (CU @ CV.T)


def make_lr_batch(ntiles, nb, k, maxrk=None, *, base_seed=1234, decades=8.0, dtype=np.float64,
                  ranks=None, first_tile=0):
    """Batch of ``ntiles`` problems with ranks ``k`` (or per-tile ``ranks[t] = (a, b, c)``).

    ``first_tile`` offsets the per-tile seed so a shard of a larger batch can be
    generated independently and still match the unsharded batch bit for bit.
    """
    maxrk = 2 * k if maxrk is None else maxrk
    au = np.zeros((ntiles, k, nb), dtype); av = np.zeros_like(au)
    bu = np.zeros((ntiles, k, nb), dtype); bv = np.zeros_like(bu)
    cu = np.zeros((ntiles, maxrk, nb), dtype); cv = np.zeros_like(cu)
    ark = np.zeros(ntiles, np.int32); brk = np.zeros(ntiles, np.int32); crk = np.zeros(ntiles, np.int32)
    real = np.float32 if dtype in (np.float32, np.complex64) else np.float64
    for t in range(ntiles):
        rng = np.random.default_rng(base_seed + first_tile + t)
        a, b, c = (k, k, k) if ranks is None else ranks[t]
        ark[t], brk[t], crk[t] = a, b, c
        for (u, v, r) in ((au, av, a), (bu, bv, b), (cu, cv, c)):
            sig = (10.0 ** (-decades * np.arange(r) / max(k, 1))).astype(real)
            u[t, :r] = orth_factor(rng, nb, r, dtype)
            v[t, :r] = orth_factor(rng, nb, r, dtype) * sig[:, None]
    return TlrBatch(nb, maxrk, au, av, bu, bv, cu, cv, ark, brk, crk)


def make_dense_batch(ntiles, nb, *, base_seed=4321, dtype=np.float64, lower_only=False, first_tile=0):
    """Dense ``nb x nb`` tiles (column-major as (nb, nb)[col, row]) for dgemmcd / dsyrk."""
    cd = np.zeros((ntiles, nb, nb), dtype)
    for t in range(ntiles):
        rng = np.random.default_rng(base_seed + first_tile + t)
        x = (rng.random((nb, nb)) - 0.5).astype(dtype)
        if lower_only:
            x = np.triu(x)  # [col, row] storage: keeping col <= row is the lower triangle
        cd[t] = x
    return cd


def _morton_points(n, rng):
    """``n`` jittered points of the unit square in Morton (Z-curve) order, as HiCMA's
    synthetic covariance problems order them (neighbouring indices = neighbouring points)."""
    side = int(np.ceil(np.sqrt(n)))
    ix, iy = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    ix, iy = ix.ravel(), iy.ravel()

    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for bit in range(16):
            out |= ((v >> np.uint64(bit)) & np.uint64(1)) << np.uint64(2 * bit)
        return out

    order = np.argsort(spread(ix) | (spread(iy) << np.uint64(1)), kind="stable")[:n]
    jit = (rng.random((n, 2)) - 0.5) * 0.4
    return (np.stack([ix[order], iy[order]], axis=1) + 0.5 + jit) / side


def compress_tile(a, acc, maxrk):
    """Dense tile -> (U, V) with orthonormal U and V carrying the singular values, truncated at the
    absolute accuracy ``acc`` (what the reference's tests do with dgesvd,
    ``testing/testing_dgemm.c:119-236``).  Returns column-major (r, nb) arrays."""
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    r = int(np.searchsorted(-s, -acc))          # first index with s < acc
    r = max(1, min(r, maxrk))
    return np.ascontiguousarray(u[:, :r].T), np.ascontiguousarray(vt[:r] * s[:r, None]), r


def make_covariance_batch(ntiles, nb, maxrk, *, acc=1e-8, base_seed=99, length=0.4, first_tile=0):
    """The realistic regime: per problem three tiles cut from ONE squared-exponential
    covariance matrix ``K(x, y) = exp(-(|x - y| / length)^2)`` (HiCMA's synthetic ``sqexp`` problem) over
    3*nb Morton-ordered points of the unit square, compressed to accuracy ``acc``: ``A = K[1, 0]``, ``B = K[2, 0]``, ``C = K[2, 1]`` -- the
    operands of the TLR-Cholesky trailing update ``C <- C - A B^T`` in which the update shares most of
    its column space with C, so the rank after recompression stays close to the ranks going in.
    Per-tile ranks differ (ragged batch); factors are zero padded to ``maxrk`` columns."""
    au = np.zeros((ntiles, maxrk, nb)); av = np.zeros_like(au)
    bu = np.zeros_like(au); bv = np.zeros_like(au)
    cu = np.zeros_like(au); cv = np.zeros_like(au)
    ark = np.zeros(ntiles, np.int32); brk = np.zeros(ntiles, np.int32); crk = np.zeros(ntiles, np.int32)
    for t in range(ntiles):
        rng = np.random.default_rng(base_seed + first_tile + t)
        pts = _morton_points(3 * nb, rng)
        d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
        k = np.exp(-(d / (length * (1.0 + 0.5 * ((first_tile + t) % 3)))) ** 2)
        blk = lambda i, j: k[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb]
        # B enters the product transposed: A B^T = K10 (K20)^T
        for (uu, vv, rr, tile) in ((au, av, ark, blk(1, 0)), (bu, bv, brk, blk(2, 0)), (cu, cv, crk, blk(2, 1))):
            u, v, r = compress_tile(tile, acc, maxrk // 2)
            uu[t, :r], vv[t, :r], rr[t] = u, v, r
    return TlrBatch(nb, maxrk, au, av, bu, bv, cu, cv, ark, brk, crk)
