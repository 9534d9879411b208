"""CPU tests of the oracle (the parity reference): the reference's own behavioural tests restated
(tests/testthat/test-sparseEigen.R), the README statistical known-answers, golden regression vectors
and internal identities.  No GPU needed."""
import numpy as np
import pytest

from oracle import speigen_oracle as o
from conftest import colwise_inner


def _rand(m=10, n=15, seed=0):
    return np.random.default_rng(seed).standard_normal((n, m))


# ---- tests/testthat/test-sparseEigen.R:1-18 "NA check" -------------------------------------------
def test_na_check():
    X = _rand()
    Xna = X.copy(); Xna[0, 0] = np.nan
    with pytest.raises(o.SpEigenError): o.spEigen(Xna, 3, 0.5, data=True)
    with pytest.raises(o.SpEigenError): o.spEigen(np.cov(Xna, rowvar=False), 3, 0.5)
    with pytest.raises(o.SpEigenError): o.spEigen(X, float("nan"), 0.5, data=True)
    with pytest.raises(o.SpEigenError): o.spEigen(np.cov(X, rowvar=False), None, 0.5)
    with pytest.raises(o.SpEigenError): o.spEigen(X, 3, float("nan"), data=True)
    with pytest.raises(o.SpEigenError): o.spEigen(np.cov(X, rowvar=False), 3, None)


# ---- :20-36 "positive integer q check" ------------------------------------------------------------
@pytest.mark.parametrize("q", [-1, 0, 0.5])
def test_positive_integer_q(q):
    X = _rand()
    with pytest.raises(o.SpEigenError): o.spEigen(X, q, 0.5, data=True)
    with pytest.raises(o.SpEigenError): o.spEigen(np.cov(X, rowvar=False), q, 0.5)


# ---- :38-52 "input dimension check" ---------------------------------------------------------------
def test_input_dimension():
    X = _rand()
    with pytest.raises(o.SpEigenError): o.spEigen(X[:, 0], 3, 0.5, data=True)          # m == 1
    with pytest.raises(o.SpEigenError): o.spEigen(np.atleast_2d(np.var(X[:, 0])), 3, 0.5)
    with pytest.raises(o.SpEigenError): o.spEigen(X[0, :], 3, 0.5, data=True)          # as.matrix(vector) is 10 x 1


def test_rho_must_be_positive():
    with pytest.raises(o.SpEigenError): o.spEigen(_rand(), 3, 0.0, data=True)          # R/spEigen.R:55


def test_not_symmetric_and_not_pd_and_rank():
    X = _rand()
    S = np.cov(X, rowvar=False)
    A = S.copy(); A[0, 1] += 1e-3
    with pytest.raises(o.SpEigenError, match="not symmetric"): o.spEigen(A, 3)
    with pytest.raises(o.SpEigenError, match="not PD"): o.spEigen(S - 2 * np.eye(10), 3)
    low = X[:3].T @ X[:3]                                                              # rank 3
    with pytest.raises(o.SpEigenError, match="rank"): o.spEigen(low, 5)


# ---- :55-76 "output dimension check" --------------------------------------------------------------
def test_output_dimensions():
    X = _rand()
    for args, kw in (((X, 3, 0.5), dict(data=True)), ((np.cov(X, rowvar=False), 3, 0.5), {})):
        r = o.spEigen(*args, **kw)
        assert r["vectors"].size == 30 and r["values"].size == 3 and r["standard_vectors"].size == 30


# ---- host constants --------------------------------------------------------------------------------
def test_schedule_and_defaults():
    pp, eps, tol = o.schedule()
    assert len(pp) == 11 and np.isclose(pp[0], 0.1) and np.isclose(pp[-1], 1e-7)       # R/spEigen.R:94-99
    assert np.allclose(tol, pp * 1e-2) and np.array_equal(eps, pp)
    assert np.allclose(o.default_d(3), [1.0, 0.75, 0.5]) and np.allclose(o.default_d(1), [1.0])
    c1, c2, w0 = o.stage_constants(0.1, 0.1)
    assert np.isclose(c1, np.log(11.0)) and np.isclose(c2, 0.4 * np.log(11.0)) and np.isclose(w0, 1 / (0.1 * c2))
    # weights are continuous at |v| = eps and w0 is their maximum
    w = o.weights(np.array([[0.05, 0.1, 0.1000001, 0.5]]), w0, c1, 0.1, 0.1)
    assert np.isclose(w[0, 1], w0) and np.isclose(w[0, 2], w0, rtol=1e-5) and w[0, 3] < w0


# ---- README statistical known-answer (README.md:87-89; seed-dependent only statistically) ---------
def test_readme_recovery():
    X, Vq = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    r = o.spEigen(o.sample_cov(X), 3)
    rec = colwise_inner(r["vectors"], Vq)
    std = colwise_inner(r["standard_vectors"], Vq)
    assert np.all(rec >= 0.997), rec                                                   # README: 0.9987 0.9988 0.9972
    assert np.all(rec > std)                                                           # sparse beats plain eigen()
    assert np.array_equal((r["vectors"] != 0).sum(axis=0), [50, 50, 50])               # planted support size
    for j in range(3):
        assert np.all(r["vectors"][j * 50:(j + 1) * 50, j] != 0)


# ---- golden regression vectors -------------------------------------------------------------------
@pytest.mark.parametrize("name", ["real_cov", "real_data", "cplx_cov", "cplx_data"])
def test_golden(golden, name):
    X, q = golden[f"{name}.X"], int(golden[f"{name}.q"])
    r = o.spEigen(X, q, rho=float(golden[f"{name}.rho"]), data=bool(golden[f"{name}.data"]), return_trace=True)
    assert np.allclose(r["values"], golden[f"{name}.values"], rtol=1e-12)
    F, Fg = np.array(r["trace"].F), golden[f"{name}.F"]
    n = min(len(F), len(Fg), 12)                         # prefix: later iterations sit on the round-off floor
    assert np.allclose(F[:n], Fg[:n], rtol=1e-9)
    assert np.all(colwise_inner(r["vectors"], golden[f"{name}.vectors"]) > 1 - 1e-5)
    assert np.array_equal(r["vectors"] != 0, golden[f"{name}.vectors"] != 0)


# ---- internal identities ---------------------------------------------------------------------------
def test_iterates_orthonormal_and_monotone():
    X, _ = o.spiked_samples(120, 60, 4, 0.1, seed=5)
    r = o.spEigen(o.sample_cov(X), 4, keep_iterates=True, return_trace=True)
    tr = r["trace"]
    for V in tr.iterates:
        assert np.abs(V.T @ V - np.eye(4)).max() < 1e-12
    F, st = np.array(tr.F), np.array(tr.stage)
    for k in range(1, len(F)):
        if st[k] == st[k - 1]:                            # accepted steps within a stage never decrease F (R/spEigen.R:140)
            assert F[k] * (1 + np.sign(F[k]) * 1e-9) > F[k - 1]


def test_reference_vs_direct_mode_prefix():
    X, _ = o.spiked_samples(300, 80, 3, 0.1, seed=9)
    S = o.sample_cov(X)
    a = o.spEigen(S, 3, mode="reference", return_trace=True)
    b = o.spEigen(S, 3, mode="direct", return_trace=True)
    Fa, Fb = np.array(a["trace"].F), np.array(b["trace"].F)
    assert np.allclose(Fa[:10], Fb[:10], rtol=1e-12)
    assert np.array_equal(a["vectors"] != 0, b["vectors"] != 0)
    assert np.all(colwise_inner(a["vectors"], b["vectors"]) > 1 - 1e-6)


def test_sign_equivariance():
    X, _ = o.spiked_samples(100, 50, 3, 0.1, seed=2)
    S = o.sample_cov(X)
    prob = o.setup(S, 3)
    a = o.spEigen(S, 3, V=prob.V)
    b = o.spEigen(S, 3, V=prob.V * np.array([1, -1, 1])[None, :])
    assert np.all(colwise_inner(a["vectors"], b["vectors"]) > 1 - 1e-6)


def test_complex_cov_and_data_agree():
    Xc, Vc = o.spiked_samples(80, 120, 2, 0.2, seed=1, cplx=True)
    r1 = o.spEigen(Xc, 2, data=True)
    assert r1["vectors"].dtype == np.complex128 and r1["values"].dtype == np.float64
    G = np.conj(r1["vectors"].T) @ r1["vectors"]
    assert np.abs(G - np.eye(2)).max() < 1e-6
