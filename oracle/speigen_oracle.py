"""CPU oracle for sparseEigen::spEigen()  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Line-by-line NumPy restatement of the reference R implementation
(/root/reference/R/spEigen.R:46-189).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the shipped
path (sparseeigen_b200/) never does.

PARITY STATUS: "parity unpinned".  The reference holds no numeric golden vectors
(tests/testthat/test-sparseEigen.R pins only error raising :9-51 and output lengths
:62-70) and R is not installed in this image, so the reference itself cannot be executed
here.  What pins this oracle instead (tests/test_oracle.py):
  * the reference's own behavioural tests restated (errors, output shapes);
  * the statistical known-answers of README.md:87-89 (sparse recovery >= 0.997, support =
    planted block) on the same spiked model;
  * internal identities: orthonormal iterates, monotone objective within a stage,
    reference-mode (G through Vx, R/spEigen.R:177) vs direct-mode (G = S@V) agreement on
    the pre-noise-floor prefix of the trajectory.

The arithmetic delegated by the reference to base R's BLAS/LAPACK (`%*%`, `svd`, `eigen`;
call sites R/spEigen.R:60,68,76,130,136,138,177,180; DESCRIPTION:26,30 pins no version) is
delegated here to NumPy/SciPy's OpenBLAS/LAPACK (gesdd / syevd|heevd).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

MAX_ITER = 1000          # R/spEigen.R:47
K_STAGES = 10            # R/spEigen.R:94
P1, PK = 1.0, 7.0        # R/spEigen.R:95-96
BACKTRACK_SLACK = 1e-9   # R/spEigen.R:140
RANK_TOL = 1e-9          # R/spEigen.R:69,79
PD_TOL = -1e-10          # R/spEigen.R:77
# The reference's backtracking `while(1)` (R/spEigen.R:126-146) has no exit other than
# acceptance; a -> -1 geometrically, and at a == -1 the candidate is the plain double MM
# step V2.  We cap the loop (documented divergence; never reached in any test).
MAX_BACKTRACKS = 200


class SpEigenError(ValueError):
    """Raised where the reference calls stop() (R/spEigen.R:52-55,69,75,77,79)."""


def h(x: np.ndarray) -> np.ndarray:
    """Hermitian transpose, R/spEigen.R:187-189."""
    return np.conj(x.T)


def schedule():
    """(p, eps, tol) continuation schedule, R/spEigen.R:94-102."""
    gamma = (PK / P1) ** (1.0 / K_STAGES)
    pp = P1 * gamma ** np.arange(0, K_STAGES + 1, dtype=np.float64)
    pp = 10.0 ** (-pp)
    return pp, pp.copy(), pp * 1e-2


def stage_constants(p: float, epsi: float):
    """c1, c2, w0 of R/spEigen.R:110-112."""
    c1 = np.log(1.0 + 1.0 / p)
    c2 = 2.0 * (p + epsi) * c1
    w0 = 1.0 / (epsi * c2)
    return c1, c2, w0


def default_d(q: int) -> np.ndarray:
    """seq(from = 1, to = 0.5, length.out = q), R/spEigen.R:64 (q == 1 gives 1)."""
    if q == 1:
        return np.array([1.0])
    return 1.0 - 0.5 * np.arange(q, dtype=np.float64) / (q - 1)


def polar(A: np.ndarray) -> np.ndarray:
    """s <- svd(A); s$u %*% h(s$v)   (R/spEigen.R:130-131,180-182)."""
    u, _, vh = np.linalg.svd(A, full_matrices=False)
    return u @ vh


def weights(V: np.ndarray, w0: float, c1: float, p: float, epsi: float) -> np.ndarray:
    """R/spEigen.R:170-173."""
    absV = np.abs(V)
    w = np.full(V.shape, w0, dtype=np.float64)
    ind = absV > epsi
    a = absV[ind]
    w[ind] = (0.5 / c1) / (a * a + p * a)
    return w


def reweight_H(V, rho_vec, w0, c1, p, epsi):
    """H of R/spEigen.R:176: (w - colmax w) * V * rho (per column)."""
    w = weights(V, w0, c1, p, epsi)
    return (w - w.max(axis=0, keepdims=True)) * V * rho_vec[None, :]


def penalty(V0, c1, c2, p, epsi):
    """g of R/spEigen.R:133-134."""
    a = np.abs(V0)
    g = np.empty(a.shape, dtype=np.float64)
    small = a <= epsi
    g[small] = a[small] ** 2 / (epsi * c2)
    big = ~small
    g[big] = np.log((p + a[big]) / (p + epsi)) / c1 + epsi / c2
    return g


@dataclasses.dataclass
class Problem:
    """Everything the MM loop needs after the init block R/spEigen.R:59-87."""
    X: np.ndarray            # (centred) data matrix or covariance
    data: bool
    q: int
    d: np.ndarray
    rho_vec: np.ndarray
    Vx: Optional[np.ndarray]  # eigen/singular vectors (None in direct mode w/o full EVD)
    sv2: np.ndarray           # eigenvalues (cov) / squared singular values (data)
    V: np.ndarray             # start point
    nrow: int
    mode: str                 # "reference" | "direct"

    def contract(self, U: np.ndarray) -> np.ndarray:
        """S+ @ U (cov) or X^H X @ U (data) *without* the column weights d.

        reference mode follows R/spEigen.R:177: Vx %*% ((h(Vx) %*% U) * sv2);
        direct mode is the contraction the B200 path performs."""
        if self.mode == "reference":
            return self.Vx @ ((h(self.Vx) @ U) * self.sv2[:, None])
        if self.data:
            return h(self.X) @ (self.X @ U)
        return self.X @ U

    def quad(self, V0: np.ndarray) -> np.ndarray:
        """Per-column quadratic term of the objective, R/spEigen.R:136,138 (always via X)."""
        if self.data:
            return np.sum(np.abs(self.X @ V0) ** 2, axis=0)
        return np.real(np.sum(np.conj(V0) * (self.X @ V0), axis=0))


def validate(X, q, rho):
    """R/spEigen.R:50-55."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 1)          # as.matrix(vector) -> n x 1
    if X.ndim != 2:
        raise SpEigenError("X must be a matrix")
    m = X.shape[1]
    if m == 1:
        raise SpEigenError("Data is univariate!")
    def _isna(v):
        return v is None or (isinstance(v, float) and np.isnan(v))
    if np.isnan(X).any() or _isna(q) or _isna(rho):
        raise SpEigenError("This function cannot handle NAs.")
    if (q % 1) != 0 or q <= 0:
        raise SpEigenError("The input argument q should be a positive integer.")
    if rho <= 0:
        raise SpEigenError("The input argument rho should be positive.")
    return X, int(q)


def is_symmetric(X: np.ndarray) -> bool:
    """isSymmetric.matrix(X), R/spEigen.R:75: square and all.equal(X, Conj(t(X)),
    tolerance = 100*.Machine$double.eps) i.e. mean relative difference."""
    if X.shape[0] != X.shape[1]:
        return False
    diff = np.abs(X - h(X)).sum()
    scale = np.abs(X).sum()
    tol = 100 * np.finfo(np.float64).eps
    if scale == 0:
        return diff == 0
    return bool(diff / scale <= tol) if np.isfinite(scale) else False


def setup(X, q=1, rho=0.5, data=False, d=None, V=None, mode="reference") -> Problem:
    """Init block R/spEigen.R:50-87."""
    X, q = validate(X, q, rho)
    cplx = np.iscomplexobj(X)
    X = X.astype(np.complex128 if cplx else np.float64)
    nrow = X.shape[0]
    if data:                                    # :59-60
        X = X - X.mean(axis=0, keepdims=True)
    d = default_d(q) if d is None else np.asarray(d, dtype=np.float64)   # :63-64
    if data:                                    # :67-73
        _, sv, vh = np.linalg.svd(X, full_matrices=False)
        if q > int(np.sum(sv > RANK_TOL)):
            raise SpEigenError("The number of estimated eigenvectors q should not be larger than rank(X).")
        sv2 = sv ** 2
        Vx = h(vh)
        rho_vec = rho * np.max(np.sum(np.abs(X) ** 2, axis=0)) * (sv2[:q] / sv2[0]) * d / d[0]
    else:                                       # :74-83
        if not is_symmetric(X):
            raise SpEigenError("The covariance matrix is not symmetric.")
        # R's eigen(symmetric=TRUE) reads the lower triangle (LAPACK uplo="L").
        vals, vecs = np.linalg.eigh(X, UPLO="L")
        vals, vecs = vals[::-1].copy(), vecs[:, ::-1].copy()   # decreasing, as eigen()
        if np.any(vals < PD_TOL):
            raise SpEigenError("The covariance matrix is not PD.")
        vals[vals < 0] = 0.0
        if q > int(np.sum(vals > RANK_TOL)):
            raise SpEigenError("The number of estimated eigenvectors q should not be larger than rank(X).")
        sv2, Vx = vals, vecs
        rho_vec = rho * np.max(np.real(np.diag(X))) * (sv2[:q] / sv2[0]) * d / d[0]
    V0 = Vx[:, :q].copy() if V is None else np.asarray(V).astype(X.dtype)   # :86-87
    return Problem(X=X, data=bool(data), q=q, d=d, rho_vec=np.asarray(rho_vec, dtype=np.float64),
                   Vx=Vx, sv2=sv2, V=V0, nrow=nrow, mode=mode)


def mm_update(prob: Problem, V, w0, c1, p, epsi):
    """spEigenMMupdate, R/spEigen.R:167-183."""
    H = reweight_H(V, prob.rho_vec, w0, c1, p, epsi)
    G = prob.contract(V * prob.d[None, :])
    return polar(G - H)


def objective(prob: Problem, V0, c1, c2, p, epsi) -> float:
    """F_v[k], R/spEigen.R:133-138."""
    g = penalty(V0, c1, c2, p, epsi)
    return float(prob.quad(V0) @ prob.d - g.sum(axis=0) @ prob.rho_vec)


@dataclasses.dataclass
class Trace:
    stage: list
    F: list
    a_init: list
    a_final: list
    backtracks: list
    nR: list
    nU: list
    iterates: list            # accepted V per k (only if keep_iterates)

    @property
    def k(self):
        return len(self.F)


def run_loop(prob: Problem, thres=1e-9, keep_iterates=False, max_iter=MAX_ITER):
    """MM loop R/spEigen.R:105-160.  Returns (vectors, Trace)."""
    pp, Eps, tol = schedule()
    V = prob.V.copy()
    tr = Trace([], [], [], [], [], [], [], [])
    F_v = []
    k = 0
    for ee in range(K_STAGES + 1):                          # :107
        p, epsi = float(pp[ee]), float(Eps[ee])
        c1, c2, w0 = stage_constants(p, epsi)
        flg = 1
        while True:                                          # :115
            k += 1
            V1 = mm_update(prob, V, w0, c1, p, epsi)         # :119
            V2 = mm_update(prob, V1, w0, c1, p, epsi)        # :120
            R = V1 - V
            U = V2 - V1 - R
            nR = np.linalg.norm(np.abs(R))
            nU = np.linalg.norm(np.abs(U))
            with np.errstate(divide="ignore", invalid="ignore"):
                a = min(-nR / nU, -1.0) if nU > 0 else -np.inf   # :123 (nU==0 -> -Inf in R)
            a0 = a
            nbt = 0
            while True:                                      # :126
                V0 = polar(V - 2 * a * R + a * a * U)        # :127-131
                Fk = objective(prob, V0, c1, c2, p, epsi)    # :133-138
                Fprev = F_v[max(k - 1, 1) - 1] if F_v else Fk
                if flg == 0 and Fk * (1 + np.sign(Fk) * BACKTRACK_SLACK) <= Fprev and nbt < MAX_BACKTRACKS:
                    a = (a - 1) / 2                          # :141
                    nbt += 1
                else:
                    V = V0
                    break
            F_v.append(Fk)
            tr.stage.append(ee + 1); tr.F.append(Fk); tr.a_init.append(a0); tr.a_final.append(a)
            tr.backtracks.append(nbt); tr.nR.append(nR); tr.nU.append(nU)
            if keep_iterates:
                tr.iterates.append(V.copy())
            if flg == 0:                                     # :149-153
                rel_change = abs(F_v[k - 1] - F_v[k - 2]) / max(1.0, abs(F_v[k - 2]))
                if rel_change <= tol[ee] or k >= max_iter:
                    break
            flg = 0
    return finalize(V, thres), tr


def finalize(V, thres):
    """Threshold and "normalise", R/spEigen.R:158-160."""
    Vout = V.copy()
    Vout[np.abs(Vout) < thres] = 0                           # :158 (strict <)
    nrm = 1.0 / np.sqrt(np.sum(np.abs(Vout) ** 2, axis=0))   # :159
    # :160 is `matrix(rep(nrm, m), ncol = q) * V`: rep() cycles the q factors through the column-major layout, so entry
    # (i, j) is scaled by nrm[(i + j*m) mod q] -- what R computes, not a per-column normalisation (they coincide to
    # ~1e-14 at the default thres, where every factor is 1 + O(m * thres^2)).
    mm, qq = Vout.shape
    lin = np.arange(mm)[:, None] + mm * np.arange(qq)[None, :]
    return Vout * nrm[lin % qq]                              # :160


def spEigen(X, q=1, rho=0.5, data=False, d=None, V=None, thres=1e-9, mode="reference",
            keep_iterates=False, return_trace=False, max_iter=MAX_ITER):
    """spEigen(X, q, rho, data, d, V, thres), R/spEigen.R:46-163.

    Returns dict(vectors, standard_vectors, values) (+ 'trace', 'problem' if asked)."""
    prob = setup(X, q, rho, data, d, V, mode=mode)
    vectors, tr = run_loop(prob, thres=thres, keep_iterates=keep_iterates, max_iter=max_iter)
    q = prob.q
    values = prob.sv2[:q] / (prob.nrow - 1) if prob.data else prob.sv2[:q].copy()   # :162
    out = dict(vectors=vectors, standard_vectors=prob.Vx[:, :q].copy(), values=values)
    if return_trace:
        out["trace"] = tr
        out["problem"] = prob
    return out


# ----------------------------------------------------------------------------------------------
# Synthetic spiked model (README.md:56-71; complex: vignettes/SparseEigenvectors-vignette.Rmd:200-217)
# ----------------------------------------------------------------------------------------------
def spiked_truth(m: int, q: int, card: float, cplx=False, rng=None):
    """q non-overlapping sparse eigenvectors of cardinality card*m, entries 1/sqrt(card*m)
    (complex: unit-modulus random phase), README.md:62-63."""
    sp = int(round(card * m))
    assert q * sp <= m
    Vq = np.zeros((m, q), dtype=np.complex128 if cplx else np.float64)
    for j in range(q):
        if cplx:
            th = rng.uniform(0, 2 * np.pi, sp)
            Vq[j * sp:(j + 1) * sp, j] = np.exp(1j * th) / np.sqrt(sp)
        else:
            Vq[j * sp:(j + 1) * sp, j] = 1.0 / np.sqrt(sp)
    lam = 100.0 * np.arange(q, 0, -1, dtype=np.float64)       # README.md:70
    return Vq, lam


def spiked_samples(m, n, q, card, seed=42, cplx=False):
    """n samples ~ N(0, R), R = I + Vq (Lam - I) Vq^H, drawn as X = Z + (Z Vq)(Lam^{1/2}-I)Vq^H
    (identical in distribution to MASS::mvrnorm(n, 0, R) of README.md:77, no m x m QR needed).
    Philox counter-based generator, fixed seed."""
    rng = np.random.Generator(np.random.Philox(seed))
    Vq, lam = spiked_truth(m, q, card, cplx, rng)
    if cplx:
        Z = (rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))) / np.sqrt(2.0)
    else:
        Z = rng.standard_normal((n, m))
    X = Z + ((Z @ Vq) * (np.sqrt(lam) - 1.0)[None, :]) @ h(Vq)
    return X, Vq


def sample_cov(X):
    """cov(X) for real X; for complex the vignette's S = 1/(n-1) t(Xc) %*% Conj(Xc)
    (vignettes/...Rmd:224), exactly Hermitian by construction after symmetrisation."""
    n = X.shape[0]
    Xc = X - X.mean(axis=0, keepdims=True)
    if np.iscomplexobj(X):
        S = (Xc.T @ np.conj(Xc)) / (n - 1)
    else:
        S = (Xc.T @ Xc) / (n - 1)
    return (S + h(S)) / 2
