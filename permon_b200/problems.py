"""Workload generators: the reference's tutorial problems and the synthetic configurations C1..C5.

Host-side numpy only (no oracle, no GPU).  Everything synthetic is derived from a counter-based RNG,
``splitmix64(seed ^ global_index)``, so any row partition generates identical data.

* ``tutorial_ex1/ex2``  -- src/tutorials/ex1.c:58-106, ex2.c:48-112 of the reference
* ``jbearing2``         -- src/tutorials/jbearing2.c:199-234 (ComputeB), :355-458 (FormHessian)
* ``obstacle2d``        -- C1 / C2 (5-point Laplacian, lower obstacle)
* ``obstacle3d``        -- C3 (7-point Laplacian, lower obstacle), row-block generation for slab partitions
* ``varcoef3d``         -- C5 (variable-coefficient 7-point, manufactured KKT, one finite bound per dof)
* ``svm_dual``          -- C4 (dual SVM, Hessian Z Z^T with Z = diag(y) X, box [0,C], equality y^T a = 0)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED = 20260101
PETSC_INFINITY = 1.7976931348623157e308 / 4.0
PETSC_NINFINITY = -PETSC_INFINITY


def splitmix64(idx, seed=SEED):
    """Vectorised splitmix64 finaliser of (seed ^ idx) -> uint64."""
    z = (np.asarray(idx, dtype=np.uint64) ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def u01(idx, stream=0, seed=SEED):
    """Uniform double in [0,1) for counter ``idx`` on independent ``stream``."""
    z = splitmix64(np.asarray(idx, dtype=np.uint64) * np.uint64(8) + np.uint64(stream), seed)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class QPProblem:
    """min 1/2 x'Ax - b'x  s.t. lb <= x <= ub [, B x = c]; rows [r0, r1) of a global N-row problem."""
    name: str
    N: int                      # global number of rows
    r0: int                     # first local row
    r1: int                     # one past the last local row
    ia: np.ndarray              # local CSR, GLOBAL column indices
    ja: np.ndarray
    a: np.ndarray
    b: np.ndarray
    lb: np.ndarray | None
    ub: np.ndarray | None
    x0: np.ndarray
    is_: np.ndarray | None = None   # bounds apply to these (global) indices only (ex2)
    B: np.ndarray | None = None     # dense equality rows (m x n_local)
    c: np.ndarray | None = None
    second: tuple | None = None     # (ia2, ja2, a2): Hessian = CSR(ia,ja,a) * CSR(second)
    meta: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.r1 - self.r0

    @property
    def nnz(self):
        return int(self.ia[-1])


def _csr_from_candidates(cols, vals, mask):
    """cols/vals/mask: (nrows, k) candidate entries in ascending column order; keep the masked ones."""
    counts = mask.sum(axis=1)
    ia = np.zeros(cols.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=ia[1:])
    assert ia[-1] < 2**31
    return ia.astype(np.int32), cols[mask].astype(np.int32), vals[mask].astype(np.float64)


# ---------------------------------------------------------------------------------------------------
# reference tutorials
# ---------------------------------------------------------------------------------------------------

def _fobst(i, n):
    h = 1.0 / (n - 1)
    return np.sin(4 * np.pi * i * h - np.pi / 6.0) / 2 - 2


def _tutorial_1d_matrix(n):
    """tridiag(-1,2,-1) with identity rows 0 and n-1 and the couplings to them dropped (ex1.c:75-100)."""
    ia, ja, a = [0], [], []
    for i in range(n):
        if i == 0 or i == n - 1:
            ja += [i]
            a += [1.0]
        else:
            if i != 1:
                ja.append(i - 1)
                a.append(-1.0)
            ja.append(i)
            a.append(2.0)
            if i != n - 2:
                ja.append(i + 1)
                a.append(-1.0)
        ia.append(len(ja))
    return np.array(ia, np.int32), np.array(ja, np.int32), np.array(a, np.float64)


def tutorial_ex1(n=100):
    h = 1.0 / (n - 1)
    ia, ja, a = _tutorial_1d_matrix(n)
    b = np.full(n, -15 * h * h * 2)
    b[0] = b[-1] = 0.0
    lb = _fobst(np.arange(n, dtype=np.float64), n)
    lb[0] = lb[-1] = 0.0  # c is only set on rstart..rend (ex1.c:99), the Dirichlet rows keep 0
    return QPProblem("ex1", n, 0, n, ia, ja, a, b, lb, None, np.zeros(n))


def tutorial_ex3_dual(n=100):
    """ex3.c dualised (QPTDualize, qptransform.c:909-1186, the -spd / empty-nullspace branch): the primal problem is ex1 with the
    bound written as the inequality -I x <= -c; the dual QP is  min 1/2 l'F l - l'd, l >= 0  with F = B K^+ B', d = B K^+ f - c_I,
    B = -I.  The reference applies K^+ with a sparse direct solver (MUMPS); here K^-1 is formed densely (n = 100), so F is an
    explicit dense matrix in CSR -- same mathematics, rounding of a different factorisation."""
    ia, ja, a = _tutorial_1d_matrix(n)
    K = np.zeros((n, n))
    for i in range(n):
        K[i, ja[ia[i]:ia[i + 1]]] = a[ia[i]:ia[i + 1]]
    h = 1.0 / (n - 1)
    f = np.full(n, -15 * h * h * 2)
    f[0] = f[-1] = 0.0
    c = _fobst(np.arange(n, dtype=np.float64), n)
    c[0] = c[-1] = 0.0
    cI = -c                                   # VecScale(c, -1), ex3.c:126
    Kinv = np.linalg.inv(K)
    Bm = -np.eye(n)                           # MatScale(B, -1), ex3.c:127
    F = Bm @ Kinv @ Bm.T
    d = Bm @ (Kinv @ f) - cI                  # d = B K^+ f - c (:1128-1132)
    fia = (np.arange(n + 1) * n).astype(np.int32)
    fja = np.tile(np.arange(n, dtype=np.int32), n)
    return QPProblem("ex3dual", n, 0, n, fia, fja, np.ascontiguousarray(F).ravel(), d, np.zeros(n), None, np.zeros(n),
                     meta=dict(Kinv=Kinv, f=f, B=Bm, cI=cI, K=K))


def tutorial_ex2(n=100, infinite=False):
    h = 1.0 / (n - 1)
    ia, ja, a = _tutorial_1d_matrix(n)
    b = np.full(n, -15 * h * h * 2)
    b[0] = b[-1] = 0.0
    half = n // 2
    if infinite:
        lb = np.zeros(n)
        idx = np.arange(1, n - 1)
        lb[idx] = np.where(idx < half, _fobst(idx.astype(np.float64), n), PETSC_NINFINITY)
        return QPProblem("ex2inf", n, 0, n, ia, ja, a, b, lb, None, np.zeros(n))
    lb = np.zeros(half)
    idx = np.arange(1, half)
    lb[idx] = _fobst(idx.astype(np.float64), n)
    return QPProblem("ex2", n, 0, n, ia, ja, a, b, lb, None, np.zeros(n), is_=np.arange(half, dtype=np.int32))


def jbearing2(nx, ny, ecc=0.1, bb=10.0):
    """Journal bearing (MINPACK-2 DPJB) Hessian and rhs in natural ordering row = j*nx + i."""
    hx = 2.0 * np.pi / (nx + 1.0)
    hy = 2.0 * bb / (ny + 1.0)
    hxhy = hx * hy
    hxhx = 1.0 / (hx * hx)
    hyhy = 1.0 / (hy * hy)

    def p(xi):
        t = 1.0 + ecc * np.cos(xi)
        return t * t * t

    ia, ja, a = [0], [], []
    rows = {}
    for i in range(nx):
        xi = (i + 1) * hx
        trule1 = hxhy * (p(xi) + p(xi + hx) + p(xi)) / 6.0
        trule2 = hxhy * (p(xi) + p(xi - hx) + p(xi)) / 6.0
        trule3 = hxhy * (p(xi) + p(xi + hx) + p(xi + hx)) / 6.0
        trule4 = hxhy * (p(xi) + p(xi - hx) + p(xi - hx)) / 6.0
        trule5, trule6 = trule1, trule2
        vdown = -(trule5 + trule2) * hyhy
        vleft = -hxhx * (trule2 + trule4)
        vright = -hxhx * (trule1 + trule3)
        vup = -hyhy * (trule1 + trule6)
        vmiddle = hxhx * (trule1 + trule2 + trule3 + trule4) + hyhy * (trule1 + trule2 + trule5 + trule6)
        for j in range(ny):
            row = j * nx + i
            ent = []
            if j > 0:
                ent.append((row - nx, vdown))
            if i > 0:
                ent.append((row - 1, vleft))
            ent.append((row, vmiddle))
            if i + 1 < nx:
                ent.append((row + 1, vright))
            if j + 1 < ny:
                ent.append((row + nx, vup))
            rows[row] = ent
    for r in range(nx * ny):
        for c_, v in rows[r]:
            ja.append(c_)
            a.append(v)
        ia.append(len(ja))
    n = nx * ny
    ii = np.arange(n) % nx
    Bvec = -(ecc * hx * hy) * np.sin((ii + 1) * hx)   # ComputeB
    b = -Bvec                                          # QPSetRhsPlus negates (qp.c:1305-1307)
    return QPProblem("jbearing2", n, 0, n, np.array(ia, np.int32), np.array(ja, np.int32), np.array(a), b,
                     np.zeros(n), np.full(n, 1000.0), np.zeros(n))


# ---------------------------------------------------------------------------------------------------
# synthetic configurations
# ---------------------------------------------------------------------------------------------------

def obstacle2d(N, bscale=-30.0, rows=None, scaled=False):
    """C1/C2: N x N interior grid, h = 1/(N+1), 5-point Laplacian (4,-1), b = bscale*h^2, sinusoidal obstacle.
    ``scaled``: the Hessian becomes D A D with d_i = 1 + u01(i)/2 (counter-based, partition independent): every value is
    distinct, so no tile of the packed format can be dictionary-coded (the incompressible case, 12 B per non-zero)."""
    n = N * N
    r0, r1 = (0, n) if rows is None else rows
    h = 1.0 / (N + 1)
    r = np.arange(r0, r1, dtype=np.int64)
    i = r % N
    j = r // N
    cols = np.stack([r - N, r - 1, r, r + 1, r + N], axis=1)
    mask = np.stack([j > 0, i > 0, np.ones_like(i, bool), i < N - 1, j < N - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), cols.shape)
    if scaled:
        dsc = lambda idx: 1.0 + 0.5 * u01(np.asarray(idx).astype(np.uint64), 9)
        vals = vals * dsc(r)[:, None] * dsc(np.clip(cols, 0, n - 1))
    ia, ja, a = _csr_from_candidates(cols, vals, mask)
    x = (i + 1) * h
    y = (j + 1) * h
    lb = 0.5 * np.sin(4 * np.pi * x - np.pi / 6) * np.sin(4 * np.pi * y - np.pi / 6) - 2.0
    b = np.full(r1 - r0, bscale * h * h)
    return QPProblem(f"obstacle2d_{N}", n, r0, r1, ia, ja, a, b, lb, None, np.zeros(r1 - r0), meta=dict(N=N, dim=2))


def obstacle3d(N, rows=None, Nz=None):
    """C3: N x N x Nz interior grid (Nz defaults to N), 7-point Laplacian (6,-1), b = -40 h^2."""
    Nz = N if Nz is None else Nz
    n = N * N * Nz
    r0, r1 = (0, n) if rows is None else rows
    h = 1.0 / (N + 1)
    r = np.arange(r0, r1, dtype=np.int64)
    i = r % N
    j = (r // N) % N
    k = r // (N * N)
    P = N * N
    cols = np.stack([r - P, r - N, r - 1, r, r + 1, r + N, r + P], axis=1)
    mask = np.stack([k > 0, j > 0, i > 0, np.ones_like(i, bool), i < N - 1, j < N - 1, k < Nz - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0]), cols.shape)
    ia, ja, a = _csr_from_candidates(cols, vals, mask)
    x, y, z = (i + 1) * h, (j + 1) * h, (k + 1) * h
    lb = 0.5 * np.sin(4 * np.pi * x - np.pi / 6) * np.sin(4 * np.pi * y - np.pi / 6) * np.sin(4 * np.pi * z - np.pi / 6) - 2.0
    b = np.full(r1 - r0, -40.0 * h * h)
    return QPProblem(f"obstacle3d_{N}x{N}x{Nz}", n, r0, r1, ia, ja, a, b, lb, None, np.zeros(r1 - r0),
                     meta=dict(N=N, Nz=Nz, dim=3))


def _varcoef_matrix(N, r):
    """rows ``r`` (global ids) of the variable-coefficient 7-point finite-volume operator."""
    P = N * N
    i = r % N
    j = (r // N) % N
    k = r // P

    def kc(idx):
        return np.where((splitmix64(idx * np.uint64(8) + np.uint64(7)) & np.uint64(1)) == 1, 1.0e4, 1.0)

    k0 = kc(r.astype(np.uint64))
    offs = [-P, -N, -1, 1, N, P]
    inside = [k > 0, j > 0, i > 0, i < N - 1, j < N - 1, k < N - 1]
    cond = []
    for off, ins in zip(offs, inside):
        nb = np.where(ins, r + off, r)
        kn = kc(nb.astype(np.uint64))
        cface = np.where(ins, 2.0 * k0 * kn / (k0 + kn), k0)  # harmonic mean; boundary face: own k
        cond.append(cface)
    diag = cond[0] + cond[1] + cond[2] + cond[3] + cond[4] + cond[5]
    cols = np.stack([r - P, r - N, r - 1, r, r + 1, r + N, r + P], axis=1)
    vals = np.stack([-cond[0], -cond[1], -cond[2], diag, -cond[3], -cond[4], -cond[5]], axis=1)
    mask = np.stack([inside[0], inside[1], inside[2], np.ones_like(i, bool), inside[3], inside[4], inside[5]], axis=1)
    return cols, vals, mask


def _varcoef_state(r):
    """manufactured solution, multipliers and bounds for global rows r."""
    ru = r.astype(np.uint64)
    xhat = 2.0 * u01(ru, 1) - 1.0
    s = u01(ru, 2)
    lam_mag = 0.01 + 0.99 * u01(ru, 3)
    dist = 0.1 + 0.9 * u01(ru, 4)
    side = u01(ru, 5) < 0.5
    lower_act = s < 0.25
    upper_act = (s >= 0.25) & (s < 0.5)
    free = s >= 0.5
    lb = np.full(r.shape, PETSC_NINFINITY)
    ub = np.full(r.shape, PETSC_INFINITY)
    lam = np.zeros(r.shape)
    lb[lower_act] = xhat[lower_act]
    lam[lower_act] = lam_mag[lower_act]
    ub[upper_act] = xhat[upper_act]
    lam[upper_act] = -lam_mag[upper_act]
    fl = free & side
    fu = free & ~side
    lb[fl] = xhat[fl] - dist[fl]
    ub[fu] = xhat[fu] + dist[fu]
    return xhat, lam, lb, ub, lower_act, upper_act


def varcoef3d(N, rows=None):
    """C5: b = A xhat - lambda so that xhat is the exact solution with a known active set."""
    n = N ** 3
    r0, r1 = (0, n) if rows is None else rows
    r = np.arange(r0, r1, dtype=np.int64)
    cols, vals, mask = _varcoef_matrix(N, r)
    ia, ja, a = _csr_from_candidates(cols, vals, mask)
    xhat, lam, lb, ub, la, ua = _varcoef_state(r)
    # A xhat needs xhat at neighbour rows (possibly outside [r0,r1)): evaluate by counter
    nbr = np.where(mask, cols, r[:, None])
    xh_n = 2.0 * u01(nbr.astype(np.uint64), 1) - 1.0
    Axh = np.where(mask, vals * xh_n, 0.0).sum(axis=1)
    b = Axh - lam
    return QPProblem(f"varcoef3d_{N}", n, r0, r1, ia, ja, a, b, lb, ub, np.zeros(r1 - r0),
                     meta=dict(N=N, dim=3, xhat=xhat, lower_active=la, upper_active=ua))


def svm_dual(n, d=None, nnz_per_row=None, C=1.0):
    """C4: dual SVM.  X (n x d) has exactly ``nnz_per_row`` entries per row, one per column stratum
    (distinct, sorted), values N(0,1)/sqrt(nnz_per_row); y = +-1; Hessian A = Z Z^T, Z = diag(y) X."""
    import scipy.sparse as sp

    d = n if d is None else d
    k = max(1, d // 1000) if nnz_per_row is None else nnz_per_row
    width = d // k
    rr = np.repeat(np.arange(n, dtype=np.uint64), k)
    tt = np.tile(np.arange(k, dtype=np.uint64), n)
    cnt = rr * np.uint64(k) + tt
    cols = (tt * np.uint64(width) + splitmix64(cnt * np.uint64(8) + np.uint64(1)) % np.uint64(width)).astype(np.int32)
    u1 = u01(cnt, 2)
    u2 = u01(cnt, 3)
    vals = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2 * np.pi * u2) / np.sqrt(k)
    y = np.where(u01(np.arange(n, dtype=np.uint64), 6) < 0.5, -1.0, 1.0)
    vals = vals * np.repeat(y, k)  # fold Y into X: Z = diag(y) X (exact, y = +-1)
    ia = (np.arange(n + 1, dtype=np.int64) * k).astype(np.int32)
    Z = sp.csr_matrix((vals, cols, ia), shape=(n, d))
    Zt = Z.T.tocsr()
    Zt.sort_indices()
    return QPProblem(f"svm_{n}x{d}", n, 0, n, ia, cols, vals, np.ones(n), np.zeros(n), np.full(n, C), np.zeros(n),
                     B=y.reshape(1, n).copy(), c=None,
                     second=(Zt.indptr.astype(np.int32), Zt.indices.astype(np.int32), Zt.data.astype(np.float64)),
                     meta=dict(d=d, y=y))


def svm_dual_fast(n, d=None, nnz_per_row=None, C=1.0, workers=None):
    """The same data as ``svm_dual`` (identical counters, identical values) generated stratum by stratum on a thread pool, with the transposed
    factor assembled directly (every column stratum is a contiguous block of rows of Z^T) instead of through a sparse transpose: the
    1M x 1M / 1e9-nnz instance of BASELINE config 4 takes about a minute instead of tens."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    d = n if d is None else d
    k = max(1, d // 1000) if nnz_per_row is None else nnz_per_row
    width = d // k
    rows = np.arange(n, dtype=np.uint64)
    y = np.where(u01(rows, 6) < 0.5, -1.0, 1.0)
    cols_t = np.empty((k, n), dtype=np.int32)        # [stratum][row]
    vals_t = np.empty((k, n), dtype=np.float64)
    zt_cnt = np.zeros(k * width + (d - k * width), dtype=np.int64)
    zt_rows = np.empty((k, n), dtype=np.int32)       # rows of Z^T entries, stratum blocks, sorted by (column, row)
    zt_vals = np.empty((k, n), dtype=np.float64)

    def one(t):
        cnt = rows * np.uint64(k) + np.uint64(t)
        c = (np.uint64(t) * np.uint64(width) + splitmix64(cnt * np.uint64(8) + np.uint64(1)) % np.uint64(width)).astype(np.int32)
        v = np.sqrt(-2.0 * np.log(1.0 - u01(cnt, 2))) * np.cos(2 * np.pi * u01(cnt, 3)) / np.sqrt(k) * y
        cols_t[t] = c
        vals_t[t] = v
        order = np.argsort(c, kind="stable")          # by column, rows ascending inside a column
        zt_rows[t] = order.astype(np.int32)
        zt_vals[t] = v[order]
        zt_cnt[t * width:(t + 1) * width] = np.bincount(c - t * width, minlength=width)

    with ThreadPoolExecutor(workers or os.cpu_count() or 1) as ex:
        list(ex.map(one, range(k)))
    ia = (np.arange(n + 1, dtype=np.int64) * k).astype(np.int32)
    ja = np.ascontiguousarray(cols_t.T).reshape(-1)
    a = np.ascontiguousarray(vals_t.T).reshape(-1)
    del cols_t, vals_t
    ia2 = np.concatenate([[0], np.cumsum(zt_cnt)]).astype(np.int32)
    return QPProblem(f"svm_{n}x{d}", n, 0, n, ia, ja, a, np.ones(n), np.zeros(n), np.full(n, C), np.zeros(n),
                     B=y.reshape(1, n).copy(), c=None, second=(ia2, zt_rows.reshape(-1), zt_vals.reshape(-1)), meta=dict(d=d, y=y))


def row_partition(N, size, align=1):
    """PETSc-style contiguous ownership ranges (PetscSplitOwnership): first N % size ranks get one extra row.
    ``align`` > 1 keeps block boundaries on multiples of ``align`` (whole grid lines / planes)."""
    blocks = N // align
    base, rem = divmod(blocks, size)
    starts = [0]
    for r in range(size):
        starts.append(starts[-1] + (base + (1 if r < rem else 0)) * align)
    starts[-1] = N
    return starts
