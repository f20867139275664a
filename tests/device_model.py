"""TEST INFRASTRUCTURE: numpy model of the *device* algorithm (not of the reference).

The CUDA pipeline in ``hcore_b200/csrc`` does not call LAPACK; it uses
  1. a flat-tree tall-skinny Householder QR (first row block: plain blocked
     Householder; every further row block: triangular-on-top-of-dense "TPQRT" step
     against the running R), column sub-panels of width W with compact-WY T factors,
  2. a pivoted, truncated Householder QR of the core ``T' = R_v R_u^T`` and a one-sided BLOCK Jacobi SVD of its
     ``R^T`` (4-column blocks, rotations derived from the 8 x 8 Gram matrix of a block pair), sorted, truncated with
     the reference's rule, ``W = T' U_r`` instead of accumulating V,
  3. application of the stored block reflectors to ``[Y; 0]``.
The complex path (zgemm_kernels.cuh) factors its panels with a plain blocked Householder QR (8-column blocks, zlarft
factors) and applies Q through the same factors: ``zqr_blocked`` / ``zapply_blocked`` at the end of this file.
This file restates exactly that arithmetic in numpy so that (a) the math can be
validated against the oracle on the CPU, and (b) per-stage device dumps can be
compared with it when debugging.  It is never used by the product.
"""
from __future__ import annotations

import numpy as np

W = 8       # sub-panel width (= the 8 columns a warp owns)
H = 128     # row-block height


def _house(alpha, sigma2):
    """LAPACK dlarfg semantics given alpha and the squared norm of the tail."""
    if sigma2 == 0.0:
        return alpha, 0.0, 0.0          # beta, tau, 1/(alpha-beta) (unused)
    beta = -np.copysign(np.sqrt(alpha * alpha + sigma2), alpha)
    tau = (beta - alpha) / beta
    return beta, tau, 1.0 / (alpha - beta)


def tsqr_factor(x, h=H, w=W):
    """Factor ``x`` (m x n, m >= n).  Returns dict with R (n x n), per-block reflector
    storage V[j] (rows of block j x n), T[j][s] (w x w upper) and the block row ranges."""
    x = np.array(x, dtype=np.float64, copy=True)
    m, n = x.shape
    assert m >= n
    h0 = min(m, h)
    # n > h0 (stack wider than the row block, device width 256): rows h0 .. n-1 of R are VIRTUAL zero rows
    # appended to the panel; columns >= h0 get no reflector in block 0 (their h0 entries are rows 0 .. h0-1
    # of R) and their reflectors are born in the following blocks, starting from a zero diagonal.
    nd0 = min(n, h0)                      # columns with a diagonal inside block 0
    bounds = [(0, h0)]
    r0 = h0
    while r0 < m:
        bounds.append((r0, min(m, r0 + h)))
        r0 += h
    nsp = (n + w - 1) // w
    R = np.zeros((n, n))
    Vs, Ts = [], []
    for j, (lo, hi) in enumerate(bounds):
        B = x[lo:hi]            # view, modified in place -> becomes V
        Tj = []
        for s in range(nsp):
            c0, c1 = s * w, min(n, s * w + w)
            pw = c1 - c0
            taus = np.zeros(pw)
            Tm = np.zeros((w, w))
            Vpan = np.zeros((hi - lo, pw))
            if j == 0 and c0 >= nd0:
                # identity reflector panel (tau = 0); nothing to update to the right either
                Tj.append(Tm)
                continue
            for ii in range(pw):
                i = c0 + ii
                if j == 0:
                    xi = B[i + 1:, i].copy()
                    alpha = B[i, i]
                else:
                    xi = B[:, i].copy()
                    alpha = R[i, i]
                # one batched reduction: dots of the current column tail with all panel columns
                if j == 0:
                    d = B[i + 1:, c0:c1].T @ xi
                    dprev = Vpan[i + 1:, :ii].T @ xi     # previous reflectors (tails below row i)
                else:
                    d = B[:, c0:c1].T @ xi
                    dprev = d[:ii]
                sigma2 = d[ii]
                beta, tau, scal = _house(alpha, sigma2)
                taus[ii] = tau
                if tau == 0.0:
                    v = np.zeros_like(xi)
                else:
                    v = xi * scal
                # in-panel update of the columns to the right
                for cc in range(ii + 1, pw):
                    c = c0 + cc
                    if j == 0:
                        wv = tau * (B[i, c] + d[cc] * scal)
                        B[i, c] -= wv
                        B[i + 1:, c] -= v * wv
                    else:
                        wv = tau * (R[i, c] + d[cc] * scal)
                        R[i, c] -= wv
                        B[:, c] -= v * wv
                if j == 0:
                    B[i, i] = beta
                    B[i + 1:, i] = v
                    Vpan[i, ii] = 1.0
                    Vpan[i + 1:, ii] = v
                else:
                    R[i, i] = beta
                    B[:, i] = v
                    Vpan[:, ii] = v
                # T column (larft, forward/columnwise): T[:ii, ii] = -tau * T[:ii,:ii] @ (V[:, :ii]^T v_ii)
                Tm[ii, ii] = tau
                if ii > 0 and tau != 0.0:
                    if j == 0:
                        z = Vpan[:, :ii].T @ Vpan[:, ii]
                    else:
                        z = dprev * scal
                    Tm[:ii, ii] = -tau * (Tm[:ii, :ii] @ z)
            # trailing update
            if c1 < n:
                if j == 0:
                    S = Vpan.T @ B[:, c1:]
                    S2 = Tm[:pw, :pw].T @ S
                    B[:, c1:] -= Vpan @ S2
                else:
                    S = R[c0:c1, c1:] + Vpan.T @ B[:, c1:]
                    S2 = Tm[:pw, :pw].T @ S
                    R[c0:c1, c1:] -= S2
                    B[:, c1:] -= Vpan @ S2
            Tj.append(Tm)
        if j == 0:
            R[:nd0, :] = np.triu(B[:nd0, :n])
        Vs.append(B)
        Ts.append(Tj)
    return dict(R=R, V=Vs, T=Ts, bounds=bounds, n=n, m=m, w=w)


def _vpanel(f, j, s):
    """Explicit reflector panel (rows of block j x pw) for block j, sub-panel s."""
    n, w = f["n"], f["w"]
    c0, c1 = s * w, min(n, s * w + w)
    B = f["V"][j]
    if j > 0:
        return B[:, c0:c1]
    if c0 >= B.shape[0]:                  # wide stack: identity reflector in block 0
        return np.zeros((B.shape[0], c1 - c0))
    Vp = np.tril(B[:, c0:c1], -c0 - 1 + 0)  # strictly below the global diagonal
    # np.tril with k=-c0-1 keeps element (row, cc) when row - cc >= c0 + 1, i.e. row > c0 + cc
    for cc in range(c1 - c0):
        Vp[c0 + cc, cc] = 1.0
    return Vp


def tsqr_apply(f, y):
    """Return ``Q @ [y; 0]`` (m x r) for ``y`` of shape (n, r)."""
    n, m, w = f["n"], f["m"], f["w"]
    nsp = (n + w - 1) // w
    r = y.shape[1]
    Y = np.array(y, dtype=np.float64, copy=True)
    out = np.zeros((m, r))
    for j in range(len(f["bounds"]) - 1, 0, -1):
        lo, hi = f["bounds"][j]
        Z = np.zeros((hi - lo, r))
        for s in range(nsp - 1, -1, -1):
            c0, c1 = s * w, min(n, s * w + w)
            pw = c1 - c0
            Vp = _vpanel(f, j, s)
            S = Y[c0:c1] + Vp.T @ Z
            S2 = f["T"][j][s][:pw, :pw] @ S
            Y[c0:c1] -= S2
            Z -= Vp @ S2
        out[lo:hi] = Z
    lo, hi = f["bounds"][0]
    Yf = np.zeros((hi - lo, r))
    nd0 = min(n, hi - lo)
    Yf[:nd0] = Y[:nd0]                    # rows >= nd0 of Y sit on virtual rows: their output is discarded
    for s in range(nsp - 1, -1, -1):
        pw = min(n, s * w + w) - s * w
        Vp = _vpanel(f, 0, s)
        S2 = f["T"][0][s][:pw, :pw] @ (Vp.T @ Yf)
        Yf -= Vp @ S2
    out[lo:hi] = Yf
    return out


def qrcp(a, tol2=0.0):
    """Householder QR with column pivoting as the device runs it (``qrcp_registers`` in tlr_kernels.cuh): the squared
    norms of the remaining columns (rows >= i) are recomputed EXACTLY in every step, the pivot is the largest one
    (ties: the smaller original column), and the factorisation is TRUNCATED at the first step whose remaining block
    has a squared Frobenius norm -- the sum of those norms -- below ``tol2``; the remaining block is dropped.  Any
    permutation is backward stable: it only has to be a good *preconditioner* for the Jacobi sweeps.  The device never
    moves columns (a pivot only marks its column) and numbers the columns that were never selected in column order;
    this model swaps, which only orders the dropped-block columns differently.
    Returns (R of shape (kq, n) upper trapezoidal, perm, kq) with ``a[:, perm] = Q R + O(sqrt(tol2))``."""
    a = np.array(a, dtype=np.float64, copy=True)
    n = a.shape[1]
    perm = np.arange(n)
    kq = n
    for i in range(n):
        vn = np.sum(a[i:, i:] ** 2, axis=0)
        if vn.sum() < tol2:
            kq = i
            break
        best = vn.max()
        cand = [j for j in range(n - i) if vn[j] == best]
        p = i + min(cand, key=lambda j: perm[i + j])
        if p != i:
            a[:, [i, p]] = a[:, [p, i]]
            perm[[i, p]] = perm[[p, i]]
        x = a[i:, i]
        beta, tau, scal = _house(x[0], float(x[1:] @ x[1:]))
        if tau != 0.0:
            v = np.concatenate([[1.0], x[1:] * scal])
            wv = tau * (v @ a[i:, i + 1:])
            a[i:, i + 1:] -= np.outer(v, wv)
        a[i, i] = beta
        a[i + 1:, i] = 0.0
    return np.triu(a)[:kq], perm, kq


def rr_pair(n, step, k):
    """Pair ``k`` (0..n/2-1) of round-robin step ``step`` (0..n-2), n even: closed form
    used by the kernel (no schedule table)."""
    if k == 0:
        a, b = n - 1, step
    else:
        a, b = (step + k) % (n - 1), (step - k + n - 1) % (n - 1)
    return (a, b) if a < b else (b, a)


def round_robin_pairs(n):
    return [[rr_pair(n, st, k) for k in range(n // 2)] for st in range(n - 1)]


TOL_ROT = 1e-15     # rotate a pair when |cos| above this
TOL_STOP = 1e-7     # a sweep whose largest pre-rotation |cos| is below this is the last one


def jacobi_svd_cols(t, max_sweeps=30):
    """One-sided (Hestenes) Jacobi on the columns of ``t``: t @ J = U * diag(sigma).
    Quadratic convergence: once a sweep's largest cosine is <= TOL_STOP the rotations
    of that same sweep bring every cosine to ~TOL_STOP**2, so no checking sweep is run.
    Returns (G = t @ J unsorted, sweeps)."""
    g = np.array(t, dtype=np.float64, copy=True)
    n = g.shape[1]
    assert n % 2 == 0
    steps = round_robin_pairs(n)
    sweeps = 0
    for sweep in range(max_sweeps):
        big = False
        for st in steps:
            for (p, q) in st:
                app = g[:, p] @ g[:, p]
                aqq = g[:, q] @ g[:, q]
                apq = g[:, p] @ g[:, q]
                den = np.sqrt(app * aqq)
                if den == 0.0 or abs(apq) <= TOL_ROT * den:
                    continue
                if abs(apq) > TOL_STOP * den:
                    big = True
                zeta = (aqq - app) / (2.0 * apq)
                tt = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1.0 + zeta * zeta))
                c = 1.0 / np.sqrt(1.0 + tt * tt)
                sn = c * tt
                gp = g[:, p].copy()
                g[:, p] = c * gp - sn * g[:, q]
                g[:, q] = sn * gp + c * g[:, q]
        sweeps += 1
        if not big:
            break
    return g, sweeps


def _rot_from_gram(app, aqq, apq):
    """(cos, sin, big) of the rotation that orthogonalises a column pair with squared norms app, aqq and inner
    product apq: x_p' = c x_p - s x_q, x_q' = s x_p + c x_q (the kernel's two-rsqrt formulation)."""
    den2 = app * aqq
    apq2 = apq * apq
    if not (den2 > 0.0 and apq2 > TOL_ROT * TOL_ROT * den2):
        return 1.0, 0.0, False
    big = apq2 > TOL_STOP * TOL_STOP * den2
    dd = aqq - app
    h2 = 2.0 * apq
    rh = 1.0 / np.sqrt(dd * dd + h2 * h2)
    u = 0.5 * (abs(dd) * rh + 1.0)
    ru = 1.0 / np.sqrt(u)
    return u * ru, (h2 if dd >= 0.0 else -h2) * rh * (0.5 * ru), big


def jacobi_pair_step(x, inblock):
    """One block-pair step of the device kernel (``jacobi_pair_step`` in tlr_kernels.cuh) on the m x 8 slice ``x``
    (columns 0-3 = block p, 4-7 = block q): Gram matrix, rotations derived from the Gram matrix alone (two-sided
    updates), accumulated 8 x 8 orthogonal J, x <- x J.  Returns (x J, big)."""
    g = x.T @ x
    j = np.eye(8)
    big = False
    rounds = []
    if inblock:
        rounds += [[(i, i ^ b) for i in range(8) if i < (i ^ b)] for b in (1, 2, 3)]
    rounds += [[(i, 4 + ((i + r) & 3)) for i in range(4)] for r in range(4)]
    for pairs in rounds:
        jr = np.eye(8)
        for (lo, hi) in pairs:
            c, s, b = _rot_from_gram(g[lo, lo], g[hi, hi], g[lo, hi])
            big = big or b
            jr[lo, lo] = c; jr[hi, hi] = c; jr[hi, lo] = -s; jr[lo, hi] = s
        g = jr.T @ g @ jr
        j = j @ jr
    return x @ j, big


def jacobi_svd_block4(t, max_sweeps=30):
    """Block one-sided Jacobi of the device kernel: 4-column blocks (the column count is padded to a multiple of 8
    with zero columns), round-robin over the blocks, one ``jacobi_pair_step`` per block pair and step.
    Returns (t @ J with the padding removed, sweeps)."""
    n0 = t.shape[1]
    n8 = (n0 + 7) // 8 * 8
    g = np.zeros((t.shape[0], n8))
    g[:, :n0] = t
    nb4 = n8 // 4
    steps = round_robin_pairs(nb4)
    sweeps = 0
    for sweep in range(max_sweeps):
        big = False
        for si, st in enumerate(steps):
            for (p, q) in st:
                cols = list(range(4 * p, 4 * p + 4)) + list(range(4 * q, 4 * q + 4))
                xj, b = jacobi_pair_step(g[:, cols], si == 0)
                g[:, cols] = xj
                big = big or b
        sweeps += 1
        if not big:
            break
    return g[:, :n0], sweeps


def core_svd_model(ru, rv, rk, acc):
    """SVD stage of the device pipeline: T' = R_v R_u^T, QRCP, Jacobi on the columns of
    R^T, sort, truncate, U_r = P * normalised columns, W = R_v (R_u^T U_r)."""
    from oracle.tlr_oracle import truncation_rank
    tp = rv @ ru.T
    n = tp.shape[1]
    n16 = (n + 15) // 16 * 16                      # the kernel works on the zero-padded n16 x n16 core
    tpp = np.zeros((n16, n16)); tpp[:n, :n] = tp
    tol2 = 1e-4 * acc * acc if rk == 0 else 1e-30 * float(np.max(np.sum(tpp * tpp, axis=0)))
    r_up, perm, kq = qrcp(tpp, tol2)
    lt = np.zeros((n16, n16))
    lt[:, :kq] = r_up.T
    gj, sweeps = jacobi_svd_block4(lt[:, :kq]) if kq >= 1 else (lt[:, :kq], 0)
    g = np.zeros((n16, n16)); g[:, :kq] = gj
    sig = np.sqrt(np.sum(g * g, axis=0))
    order = np.argsort(-sig, kind="stable")
    sig = sig[order][:n]
    g = g[:, order][:, :n]
    r = truncation_rank(sig, rk, acc)
    ut = np.where(sig[:r] > 0, g[:, :r] / np.where(sig[:r] > 0, sig[:r], 1.0), 0.0)
    up = np.zeros_like(ut)
    up[perm, :] = ut                      # U = P * U~
    u = up[:n]
    w = rv @ (ru.T @ u)
    return u, w, sig, r, sweeps


def tlr_gemm_model(au, av, bu, bv, cu, cv, *, alpha, beta, rk, maxrk, acc):
    """Whole device pipeline in numpy (math orientation, nb x rank inputs)."""
    from oracle.tlr_oracle import truncation_rank
    xu = np.hstack([beta * cu, alpha * au])
    p = av.T @ bv
    xv = np.hstack([cv, bu @ p.T])
    fu = tsqr_factor(xu)
    fv = tsqr_factor(xv)
    u, wv, sig, r, sweeps = core_svd_model(fu["R"], fv["R"], rk, acc)
    return tsqr_apply(fu, u), tsqr_apply(fv, wv), r, sig, sweeps


# ---- complex path: 8-column blocks with compact-WY factors (hcore_b200/csrc/zgemm_kernels.cuh) ---------------
ZB = 8      # reflector block width of ztlr_qr_blocked_kernel / ztlr_apply_wy_kernel


def _zlarfg(alpha, xn2):
    """LAPACK zlarfg given alpha and the squared norm of the tail: beta (real), tau, 1 / (alpha - beta)."""
    if xn2 == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0 + 0.0j, 0.0 + 0.0j
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    return beta, complex((beta - alpha.real) / beta, -alpha.imag / beta), 1.0 / (alpha - beta)


def _unit_lower(blk, j0):
    """The reflector vectors of a factored block: stored part below the diagonal, unit diagonal, zero above."""
    v = blk.copy()
    for c in range(v.shape[1]):
        v[: j0 + c, c] = 0.0
        v[j0 + c, c] = 1.0
    return v


def zqr_blocked(x, b=ZB):
    """Blocked Householder QR of a complex panel the way ztlr_qr_blocked_kernel does it: the b columns of a block are
    factored column by column (zgeqr2 order, H^H applied inside the block), T is built with the zlarft recurrence
    T(k,k) = tau_k, T(0:k,k) = -tau_k T(0:k,0:k) (V(:,0:k)^H v_k), and the trailing columns get
    X <- X - V (T^H (V^H X)).  Returns (factored panel, tau, list of T blocks)."""
    x = np.array(x, dtype=np.complex128)
    m, n = x.shape
    tau = np.zeros(n, dtype=np.complex128)
    tf = []
    for j0 in range(0, n, b):
        nbk = min(b, n - j0)
        for jj in range(nbk):
            j = j0 + jj
            beta, tj, scale = _zlarfg(x[j, j], float(np.vdot(x[j + 1:, j], x[j + 1:, j]).real))
            tau[j] = tj
            if tj != 0.0:
                x[j, j] = beta
            x[j + 1:, j] *= scale
            v = np.concatenate(([1.0], x[j + 1:, j]))
            for c in range(j + 1, j0 + nbk):
                x[j:, c] -= np.conj(tj) * np.vdot(v, x[j:, c]) * v
        vb = _unit_lower(x[:, j0:j0 + nbk], j0)
        s = vb.conj().T @ vb
        t = np.zeros((b, b), dtype=np.complex128)
        for k in range(nbk):
            t[k, k] = tau[j0 + k]
            t[:k, k] = -tau[j0 + k] * (t[:k, :k] @ s[:k, k])
        tf.append(t)
        if j0 + nbk < n:
            w = vb.conj().T @ x[:, j0 + nbk:]
            x[:, j0 + nbk:] -= vb @ (t[:nbk, :nbk].conj().T @ w)
    return x, tau, tf


def zapply_blocked(xf, tf, y, b=ZB):
    """Q [y; 0] the way ztlr_apply_wy_kernel does it: blocks of b reflectors, last block first, Z <- Z - V (T (V^H Z))."""
    m, n = xf.shape
    z = np.zeros((m, y.shape[1]), dtype=np.complex128)
    z[: y.shape[0]] = y
    for bk in range(len(tf) - 1, -1, -1):
        j0 = bk * b
        nbk = min(b, n - j0)
        vb = _unit_lower(xf[:, j0:j0 + nbk], j0)
        z -= vb @ (tf[bk][:nbk, :nbk] @ (vb.conj().T @ z))
    return z
