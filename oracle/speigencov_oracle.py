"""CPU oracle for sparseEigen::spEigenCov()  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Line-by-line NumPy restatement of /root/reference/R/spEigenCov.R:48-338 (driver :48-136, `h` :140-142,
`eigvecUpdate` :146-160, `eigvalAlgo` :164-338).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this module; the shipped path (sparseeigen_b200/) never does.

PARITY STATUS: "parity unpinned".  The reference's tests pin only error raising and output lengths for this function
(tests/testthat/test-sparseEigen.R:11,14,17,33-35,47,51,72-76) and R is not installed in this image.  What pins this
restatement instead (tests/test_oracle_cov.py): the restated testthat behaviours, the README statistical
known-answers (README.md:103-121: recovery >= 0.999 and a covariance error about half the sample covariance's), the
algebraic identities of the update (monotone objective within a stage, ordered eigenvalues, unitary iterate).

Quirks of the reference kept on purpose (each is what R actually computes):
  * `A` is formed once per iteration from the OLD `V`/`Xi` (:104) and reused for the eigenvalue update (:110) after
    `V` has been replaced (:107);
  * `seq(i:q)` (:215,:218) is `1..(q-i+1)`, not `i..q`;
  * the final normalisation `matrix(rep(nrm, m), ncol = m) * V` (:133) scales ROW i by `nrm[i]`;
  * q == 1 (:289-335): when no trailing value violates the order the inner loop reads an undefined `tmp` and R stops
    with "object 'tmp' not found" -> `EigvalAlgoError` here (the product path pools `{q}` alone instead).
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .speigen_oracle import SpEigenError, h, is_symmetric, stage_constants

MAX_ITER = 3000          # R/spEigenCov.R:49
K_STAGES = 5             # :80
P1, PK = 1.0, 5.0        # :81-82
PSD_TOL = -1e-10         # :64  (`<=`)
RANK_TOL = 1e-9          # :66
SHIFT = 1e-6             # :70
MAX_PASSES = 10000       # cap on eigvalAlgo's `while (1)` (:171,:290), which has no exit on some inputs


class EigvalAlgoError(RuntimeError):
    """R would stop with an evaluation error inside eigvalAlgo (undefined `tmp`)."""


def schedule():
    """pp, Eps, tol of R/spEigenCov.R:80-88."""
    gamma = (PK / P1) ** (1.0 / K_STAGES)
    pp = 10.0 ** (-(P1 * gamma ** np.arange(0, K_STAGES + 1, dtype=np.float64)))
    return pp, pp * 1e-1, pp * 1e-1


def validate(S, q, rho):
    """R/spEigenCov.R:52-60 (same order of checks)."""
    S = np.asarray(S)
    if S.ndim == 0:
        S = S.reshape(1, 1)
    if S.ndim == 1:
        S = S.reshape(-1, 1)
    m = S.shape[1]
    if m == 1:
        raise SpEigenError("Data is univariate!")

    def _isna(v):
        return v is None or (isinstance(v, float) and np.isnan(v))
    if not _isna(q) and q > m:
        raise SpEigenError("The number of sparse eigenvectors q should not be larger than m.")
    if np.isnan(S).any() or _isna(q) or _isna(rho):
        raise SpEigenError("This function cannot handle NAs.")
    if not is_symmetric(S):
        raise SpEigenError("The covariance matrix is not symmetric")
    if (q % 1) != 0 or q <= 0:
        raise SpEigenError("The input argument q should be a positive integer.")
    if rho <= 0:
        raise SpEigenError("The input argument rho should be positive.")
    return S, int(q)


def _rmean(g, idx):
    """`1/length(ind) * sum(g[ind])` as R evaluates it: sum() accumulates in long double (80-bit on x86-64), in index
    order, rounds to double, then multiplies by the reciprocal of the count."""
    acc = np.longdouble(0.0)
    for t in idx:
        acc += np.longdouble(g[t])
    return (1.0 / len(idx)) * float(acc)


def _pool(g, g_new, ind, redInd):
    """The shared inner `while (1)` of R/spEigenCov.R:221-239,257-274: grow `ind` by the smallest remaining violators
    until the pooled mean no longer exceeds any remaining one.  1-based index lists; g/g_new are 1-based arrays."""
    ind = list(ind)
    redInd = list(redInd)
    tmp = None
    while True:
        if len(redInd) > 0:
            vals = g[redInd]
            mn = vals.min()
            indMin = [t for t in range(len(redInd)) if vals[t] == mn]
            for t in indMin:
                if redInd[t] not in ind:
                    ind.append(redInd[t])             # unique(c(ind, redInd[indMin]))
            redInd = [r for t, r in enumerate(redInd) if t not in indMin]
            tmp = _rmean(g, ind)
        if len(redInd) > 0:
            if np.sum(tmp > g[redInd]) == 0:
                g_new[ind] = _rmean(g, ind)
                break
        else:
            g_new[ind] = _rmean(g, ind)
            break


def eigval_algo(g_in, x, q):
    """eigvalAlgo(g, x, q), R/spEigenCov.R:164-338.  Returns 1/phi (the new eigenvalues)."""
    m = len(g_in)
    g = np.empty(m + 1)
    g[0] = np.nan
    g[1:] = g_in                                            # 1-based
    g_new = g.copy()
    mult = 0
    if q >= m:
        raise EigvalAlgoError("q == m: g[(q+1):m] runs backwards in R (not supported)")

    def order_ok(gv):
        return np.array_equal(np.argsort(gv[1:], kind="stable")[:q] + 1, np.arange(1, q + 1))   # :281,:328

    passes = 0
    if q > 1:
        while True:                                         # :171
            passes += 1
            if passes > MAX_PASSES:
                raise EigvalAlgoError("eigvalAlgo does not terminate on this input (R would loop forever)")
            flg = 0
            i = 1
            while i <= q - 1:                               # :177
                if g[i] >= g[i + 1]:
                    if i == q - 1:
                        flg = 1
                    j = i + 1                               # :185
                    while j <= q - 1:
                        if g[j] >= g[j + 1]:
                            if j == q - 1:
                                flg = 1
                            j += 1
                        else:
                            break
                    if flg == 0:                            # :198-200
                        g_new[i:j + 1] = _rmean(g, range(i, j + 1))
                    else:                                   # :202-242
                        swapInd = [t for t in range(q + 1, m + 1) if g[q] >= g[t]]
                        if len(swapInd) == 0:
                            g_new[i:j + 1] = _rmean(g, range(i, j + 1))
                            break
                        if mult == 1:                       # :212-216
                            ind = [t for t in range(q + 1, m + 1) if g[q] == g[t]]
                            ind = list(range(1, q - i + 2)) + ind          # c(seq(i:q), ind)
                        else:
                            ind = list(range(1, q - i + 2))                # seq(i:q)  (quirk)
                        ind = list(dict.fromkeys(ind))
                        _pool(g, g_new, ind, swapInd)
                        i = j                               # :241
                elif i == q - 1 and flg == 0:               # :245-275
                    swapInd = [t for t in range(q + 1, m + 1) if g[q] >= g[t]]
                    if mult == 1:
                        ind = [q] + [t for t in range(q + 1, m + 1) if g[q] == g[t]]
                    else:
                        ind = [q]
                    _pool(g, g_new, ind, swapInd)
                i += 1
            g = g_new.copy()                                # :279
            if order_ok(g):
                break
            mult = 1
    else:                                                   # q == 1, :289-335
        while True:
            passes += 1
            if passes > MAX_PASSES:
                raise EigvalAlgoError("eigvalAlgo does not terminate on this input (R would loop forever)")
            swapInd = [t for t in range(q + 1, m + 1) if g[q] >= g[t]]
            redInd = list(swapInd)
            if mult == 1:
                ind = [q] + [t for t in range(q + 1, m + 1) if g[q] == g[t]]
            else:
                ind = [q]
            if len(redInd) > 0:
                vals = g[redInd]
                mn = vals.min()
                for t in range(len(redInd)):
                    if vals[t] == mn and redInd[t] not in ind:
                        ind.append(redInd[t])
            else:
                raise EigvalAlgoError("object 'tmp' not found (R/spEigenCov.R:312, q == 1 without violators)")
            g_new[ind] = _rmean(g, ind)            # both branches of :312-319 do this and break
            g = g_new.copy()
            if order_ok(g):
                break
            # note: the reference never sets mult <- 1 in this branch
    gv = g[1:]
    phi = (1.0 + np.sqrt(1.0 + 4.0 * x * gv)) / (2.0 * x)   # :286,:333
    return 1.0 / phi


def weights_full(V, q, w0, c1, p, epsi):
    """w of R/spEigenCov.R:147-151 as an m x m matrix (columns > q are zero)."""
    m = V.shape[0]
    absV = np.abs(V[:, :q])
    w = np.full((m, q), w0, dtype=np.float64)
    ind = absV > epsi
    a = absV[ind]
    w[ind] = (0.5 / c1) / (a * a + p * a)
    W = np.zeros((m, m))
    W[:, :q] = w
    return W


def procrustes_target(V, A, S_hat, rho_vec, q, w0, c1, p, epsi):
    """-(H1 + H2), the matrix whose polar factor is the eigenvector update (R/spEigenCov.R:147-157)."""
    m = V.shape[0]
    W = weights_full(V, q, w0, c1, p, epsi)
    rho_full = np.concatenate([rho_vec, np.zeros(m - q)])
    H1 = (W - W.max(axis=0, keepdims=True)) * V * rho_full[None, :]      # :153
    H2 = S_hat @ A                                                        # :154
    return -(H1 + H2)


def eigvec_update(V, A, S_hat, rho_vec, q, w0, c1, p, epsi):
    """eigvecUpdate, R/spEigenCov.R:146-160."""
    u, _, vh = np.linalg.svd(procrustes_target(V, A, S_hat, rho_vec, q, w0, c1, p, epsi))   # :157
    return u @ vh                                                         # :159


def penalty(Vq, c1, c2, p, epsi):
    """g of R/spEigenCov.R:114-116."""
    a = np.abs(Vq)
    g = np.empty(a.shape)
    small = a <= epsi
    g[small] = a[small] ** 2 / (epsi * c2)
    g[~small] = np.log((p + a[~small]) / (p + epsi)) / c1 + epsi / c2
    return g


@dataclasses.dataclass
class CovTrace:
    stage: list
    F: list
    Xi: list          # eigenvalues after each iteration (only if keep_iterates)
    V: list


def spEigenCov(S, q=1, rho=0.5, thres=1e-9, return_trace=False, keep_iterates=False, max_iter=MAX_ITER):
    """spEigenCov(S, q, rho, thres), R/spEigenCov.R:48-136 -> dict(vectors m x m, values m, cov m x m)."""
    S, q = validate(S, q, rho)
    cplx = np.iscomplexobj(S)
    S = S.astype(np.complex128 if cplx else np.float64)
    m = S.shape[1]
    vals, vecs = np.linalg.eigh(S, UPLO="L")                  # eigen(S), :63
    vals, vecs = vals[::-1].copy(), vecs[:, ::-1].copy()
    if np.any(vals <= PSD_TOL):
        raise SpEigenError("The covariance matrix is not PSD.")
    vals[vals < 0] = 0.0
    if int(np.sum(vals > RANK_TOL)) < m:
        raise SpEigenError("The covariance matrix is low-rank.")
    V = vecs
    Xi = vals
    Xi_max = Xi[0]
    S_hat = S - (Xi_max + SHIFT) * np.eye(m)                  # :70
    rho_vec = rho * 10.0 * np.sqrt(np.sum(np.abs(S) ** 2)) * Xi[:q] / Xi[0]     # :73
    pp, Eps, tol = schedule()
    F_v = []
    tr = CovTrace([], [], [], [])
    k = 0
    for ee in range(K_STAGES + 1):                            # :92
        p, epsi = float(pp[ee]), float(Eps[ee])
        c1, c2, w0 = stage_constants(p, epsi)
        flg = 1
        while True:
            k += 1
            A = V / Xi[None, :]                               # V %*% diag(1/Xi), :104
            V = eigvec_update(V, A, S_hat, rho_vec, q, w0, c1, p, epsi)   # :107
            G = -np.real(np.sum(np.conj(A) * (S_hat @ A), axis=0))       # Re(diag(-h(A) S_hat A)), :110-111
            Xi = eigval_algo(G, Xi_max, q)                    # :111
            g = penalty(V[:, :q], c1, c2, p, epsi)            # :114-116
            quad = np.real(np.sum(np.conj(V) * (S @ V), axis=0))
            Fk = float(np.sum(np.log(Xi)) + np.sum(quad / Xi) + rho_vec @ g.sum(axis=0))   # :118
            F_v.append(Fk)
            tr.stage.append(ee + 1)
            tr.F.append(Fk)
            if keep_iterates:
                tr.Xi.append(Xi.copy())
                tr.V.append(V.copy())
            if flg == 0:                                      # :121-126
                rel_change = abs(F_v[k - 1] - F_v[k - 2]) / max(1.0, abs(F_v[k - 2]))
                if rel_change <= tol[ee] or k >= max_iter:
                    break
            flg = 0
    Vout = V.copy()
    Vout[np.abs(Vout) < thres] = 0                            # :131
    nrm = 1.0 / np.sqrt(np.sum(np.abs(Vout) ** 2, axis=0))    # :132
    Vout = nrm[:, None] * Vout                                # :133 (row i scaled by nrm[i]: reference quirk)
    out = dict(vectors=Vout, values=Xi.copy(), cov=(Vout * Xi[None, :]) @ h(Vout))   # :135
    if return_trace:
        out["trace"] = tr
        out["rho_vec"] = rho_vec
    return out
