"""Synthetic inputs for benchmarks and tools: the spiked covariance model of the reference's README
(README.md:56-71; complex variant vignettes/SparseEigenvectors-vignette.Rmd:200-217).

population covariance R = I + Vq (Lambda - I) Vq^H, Lambda = 100 * (q, q-1, ..., 1), Vq = q non-overlapping sparse
eigenvectors of cardinality card*m; samples X = Z + (Z Vq)(Lambda^{1/2} - I) Vq^H with Z i.i.d. N(0, 1) (CN(0, 1)), which is
MASS::mvrnorm(n, 0, R) in distribution without an m x m factorisation.  Counter-based Philox stream, fixed seed."""
import numpy as np


def spiked_truth(m, q, card, cplx=False, rng=None):
    sp = int(round(card * m))
    if q * sp > m:
        raise ValueError("q * card * m must not exceed m")
    Vq = np.zeros((m, q), dtype=np.complex128 if cplx else np.float64)
    for j in range(q):
        if cplx:
            th = rng.uniform(0, 2 * np.pi, sp)
            Vq[j * sp:(j + 1) * sp, j] = np.exp(1j * th) / np.sqrt(sp)
        else:
            Vq[j * sp:(j + 1) * sp, j] = 1.0 / np.sqrt(sp)
    return Vq, 100.0 * np.arange(q, 0, -1, dtype=np.float64)


def spiked_samples(m, n, q, card, seed=42, cplx=False):
    """(X n x m, Vq m x q)."""
    rng = np.random.Generator(np.random.Philox(seed))
    Vq, lam = spiked_truth(m, q, card, cplx, rng)
    if cplx:
        Z = (rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))) / np.sqrt(2.0)
    else:
        Z = rng.standard_normal((n, m))
    X = Z + ((Z @ Vq) * (np.sqrt(lam) - 1.0)[None, :]) @ np.conj(Vq.T)
    return X, Vq
