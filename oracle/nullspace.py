"""Oracle: pseudo-inverse / null-space / tolerance-RREF helpers (float64 NumPy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
`atacom/utils/null_space_coordinate.py:8-26` (pinv_null) and `:40-79` (rref).

Two null-space bases are offered:

* `svd_null_basis`  — what the reference does: SciPy `linalg.svd` (LAPACK gesdd),
  rank cut at eps*max(M,N)*sigma_max, trailing right-singular vectors.  For a
  null space of dimension > 1 that basis is LAPACK-implementation-defined.
* `canonical_null_basis` — the basis the CUDA kernels use: Gram-Schmidt of the
  projected unit vectors P e_0, P e_1, ... in column order, skipping a column
  when what is left of it has norm <= tol.  It depends only on the projector
  P = I - Jc^+ Jc, i.e. on null(Jc) itself, not on any SVD implementation.

Feeding either basis to `tol_rref` (the reference's RREF, restated) gives the
same Nc whenever the tolerance branch does not fire ("stratum I"); when it does
fire the reference's own output depends on which orthonormal basis LAPACK
happened to return, and `tol_rref(canonical_null_basis(...))` is the member of
that family the kernels reproduce.
"""
import numpy as np
from scipy import linalg as sla

EPS64 = np.finfo(np.float64).eps


def svd_pinv_null(Jc):
    """(Jc^+, Q, rank): Moore-Penrose inverse and SVD null-space basis.

    Restates null_space_coordinate.py:8-26.  Q is N x (N - rank).
    """
    Jc = np.asarray(Jc, dtype=np.float64)
    U, sig, Vh = sla.svd(Jc, full_matrices=True, check_finite=False)
    rows, cols = U.shape[0], Vh.shape[1]
    cut = (sig.max() if sig.size else 0.0) * EPS64 * max(rows, cols)
    rank = int(np.count_nonzero(sig > cut))
    Q = Vh[rank:].T.copy()
    pinv = (Vh[:rank].T / sig[:rank]) @ U[:, :rank].T
    return pinv, Q, rank


def default_rref_tol(V):
    """Default tolerance of the reference's rref when tol=None
    (null_space_coordinate.py:48-49): max(m,n) * eps * max-abs-row-sum."""
    m, n = V.shape
    return max(m, n) * EPS64 * np.abs(V).sum(axis=1).max()


def tol_rref(V, tol=None, trace=None):
    """Reduced row echelon form of the rows of V with a pivot tolerance.

    Restates null_space_coordinate.py:40-79 for row vectors: walk the columns;
    the pivot candidate is the first largest-|.| entry among the not-yet-used
    rows; if it is <= tol those entries are zeroed and the column skipped,
    else rows are swapped (columns j.. only), the pivot row scaled, and the
    column eliminated from every other row.

    trace (optional dict) receives: 'pivots' (list of (row, col, |p|)),
    'dropped' (list of (col, rows_used_so_far, |p|)), 'tol'.
    """
    W = np.array(V, dtype=np.float64, copy=True)
    m, n = W.shape
    if tol is None:
        tol = default_rref_tol(W)
    pivots, dropped = [], []
    r = 0
    for j in range(n):
        if r >= m:
            break
        col = np.abs(W[r:, j])
        k = r + int(np.argmax(col))
        p = abs(W[k, j])
        if p <= tol:
            dropped.append((j, r, float(p)))
            W[r:, j] = 0.0
            continue
        pivots.append((r, j, float(p)))
        if k != r:
            W[[r, k], j:] = W[[k, r], j:]
        lead = W[r, j:] / W[r, j]
        W[:, j:] -= np.outer(W[:, j], lead)
        W[r, j:] = lead
        r += 1
    if trace is not None:
        trace.update(pivots=pivots, dropped=dropped, tol=float(tol))
    return W


def null_projector(Jc, pinv=None):
    Jc = np.asarray(Jc, dtype=np.float64)
    if pinv is None:
        pinv = svd_pinv_null(Jc)[0]
    P = np.eye(Jc.shape[1]) - pinv @ Jc
    return 0.5 * (P + P.T)


def canonical_null_basis(Jc, k, tol, pinv=None, trace=None):
    """Column-ordered Gram-Schmidt basis of null(Jc) (rows of the result).

    Row i is the unit vector of the still-unused part of the null space with
    the largest component along the i-th accepted column; a column whose
    largest attainable component is <= tol is skipped.  Returns V (k x N) with
    orthonormal rows in echelon form; fewer than k accepted columns (rank
    deficiency) leaves trailing zero rows.
    """
    P = null_projector(Jc, pinv)
    N = P.shape[0]
    V = np.zeros((k, N))
    piv, drop = [], []
    r = 0
    for c in range(N):
        if r >= k:
            break
        d = P[c, c]
        p = np.sqrt(d) if d > 0.0 else 0.0
        if p <= tol:
            drop.append((c, r, float(p)))
            continue
        v = P[c] / p
        V[r] = v
        piv.append((r, c, float(p)))
        P = P - np.outer(v, v)
        r += 1
    if trace is not None:
        trace.update(pivots=piv, dropped=drop, tol=float(tol))
    return V
