"""Oracle: pseudo-inverse / null-space / tolerance-RREF helpers (float64 NumPy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
`atacom/utils/null_space_coordinate.py:8-26` (pinv_null) and `:40-79` (rref).

Three null-space bases are offered:

* `svd_pinv_null`  — what the reference does: SciPy `linalg.svd` (LAPACK gesdd),
  rank cut at eps*max(M,N)*sigma_max, trailing right-singular vectors.
* `lapack_null_basis` — the SAME basis, restated: for a full-row-rank M x N
  matrix (M < N) gesdd never touches the null vectors with its bidiagonal SVD
  iteration — with JOBZ='A' the trailing N - M rows of VT are [0 I] P^T, P the
  product G_1 ... G_M of the right Householder reflectors of the
  bidiagonalisation (dgebrd; LAPACK's path 5t, taken while N < 11 M / 6) or of
  the LQ factorisation (dgelqf / dorglq; path 4t, N >= 11 M / 6).  So the
  "implementation-defined" SVD null basis is in fact a deterministic, smooth
  function of Jc: the last N - M columns of G_1 ... G_M with LAPACK's dlarfg
  sign convention (beta = -sign(alpha) * norm).  It reproduces SciPy's vectors
  to 1e-12 including their signs (tests/test_oracle.py) — and with them the
  reference's output on the stratum where the tolerance branch of rref fires.
  This is the basis the CUDA kernels reproduce by default.
* `canonical_null_basis` — the basis of the kernels' fast path: Gram-Schmidt of
  the projected unit vectors P e_0, P e_1, ... in column order, skipping a column
  when what is left of it has norm <= tol.  It depends only on the projector
  P = I - Jc^+ Jc, i.e. on null(Jc) itself.

Feeding any orthonormal basis to `tol_rref` (the reference's RREF, restated)
gives the same Nc whenever the tolerance branch does not fire ("stratum I");
when it fires the output depends on the basis.  A column can only fire in the
reference if the canonical pivot (the norm of what is left of P e_j) is at most
sqrt(k - r) * tol (r pivots found so far): the kernels run the basis-free fast
path, flag the environments inside that band, and redo those with the LAPACK
basis.
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


def _dlarfg(alpha, x):
    """LAPACK dlarfg: H = I - tau [1; v][1; v]^T with H [alpha; x] = [beta; 0].  Returns (beta, tau, v)."""
    xnorm = float(np.sqrt(np.dot(x, x)))
    if xnorm == 0.0:
        return alpha, 0.0, np.zeros_like(x)
    beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
    return beta, (beta - alpha) / beta, x / (alpha - beta)


def lapack_null_basis(Jc):
    """Null-space basis of a full-row-rank M x N matrix exactly as LAPACK gesdd (JOBZ='A') returns it in
    VT[M:, :].T: the last N - M columns of the product of the right Householder reflectors of dgebd2 (lower
    bidiagonal form: G_i annihilates row i right of the diagonal, then H_i annihilates column i below the
    subdiagonal) or, when N >= int(11 M / 6), of dgelq2 (no left reflectors).  Restates what
    null_space_coordinate.py:9 obtains from scipy.linalg.svd for the reference's shapes."""
    A = np.array(Jc, dtype=np.float64, copy=True)
    m, n = A.shape
    lq_path = n >= int(m * 11.0 / 6.0)
    P = np.eye(n)
    for i in range(m):
        beta, tau, v = _dlarfg(A[i, i], A[i, i + 1:])
        vv = np.concatenate([[1.0], v])
        A[i, i], A[i, i + 1:] = beta, 0.0
        if i + 1 < m:
            A[i + 1:, i:] -= tau * np.outer(A[i + 1:, i:] @ vv, vv)
        P[:, i:] -= tau * np.outer(P[:, i:] @ vv, vv)
        if not lq_path and i < m - 1:
            beta, tauq, u = _dlarfg(A[i + 1, i], A[i + 2:, i])
            uu = np.concatenate([[1.0], u])
            A[i + 1, i], A[i + 2:, i] = beta, 0.0
            A[i + 1:, i + 1:] -= tauq * np.outer(uu, uu @ A[i + 1:, i + 1:])
    return P[:, m:]


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
