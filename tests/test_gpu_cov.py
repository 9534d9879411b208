"""GPU parity tests (-m gpu) for SURVEY §8(f)-1 spEigenCov() and §8(f)-2 on-device cov(X): the dense FP64 tensor-core
GEMM, the block-Jacobi polar factor / eigen-decomposition and the whole spEigenCov() loop, called through the C ABI,
against the oracle (oracle/speigencov_oracle.py) and LAPACK on the same seeded inputs.

Tolerances: GEMM <= 1e-13 relative (Frobenius); eigenvalues <= 1e-12 relative to the largest; polar factor <= 50 * eps *
cond(B) against the LAPACK-svd polar factor (both are backward stable; the factor itself is only determined to
eps * cond), orthonormal to 1e-13; spEigenCov: objective trajectory <= 1e-9 relative, eigenvalues <= 1e-9, identical
support of the q sparse eigenvectors, per-column |<u_ref, u_gpu>| >= 1 - 1e-8, cov <= 1e-6 relative."""
import numpy as np
import pytest

import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
from oracle import speigencov_oracle as oc
from conftest import colwise_inner

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ---- GEMM ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 200, 150), (257, 129, 33), (64, 500, 1000), (1000, 1, 7),
                                   (513, 640, 2049)])
@pytest.mark.parametrize("ak,bk", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_gemm_layouts(M, N, K, ak, bk):
    rng = np.random.default_rng(M + 3 * N + 7 * K + ak * 2 + bk)
    opA = rng.standard_normal((M, K))
    opB = rng.standard_normal((K, N))
    A = opA.T.copy() if ak else opA            # K x M if kmajor else M x K
    B = opB if bk else opB.T.copy()            # K x N if kmajor else N x K
    C = sb.dense_gemm(A, B, bool(ak), bool(bk), alpha=0.5)
    assert rel(C, 0.5 * opA @ opB) < 1e-13


@pytest.mark.parametrize("m,n", [(100, 37), (384, 1000), (777, 300)])
def test_gemm_symmetric_mode_and_cov(m, n):
    rng = np.random.default_rng(m + n)
    X = rng.standard_normal((n, m)) + 3.0
    G = sb.dense_gemm(X, X, True, True, symmetric=True)
    assert rel(G, X.T @ X) < 1e-13 and np.array_equal(G, G.T)
    S = sb.cov(X)
    assert rel(S, np.cov(X, rowvar=False)) < 1e-12 and np.array_equal(S, S.T)


def test_gemm_is_deterministic():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((900, 700)); B = rng.standard_normal((900, 500))
    assert np.array_equal(sb.dense_gemm(A, B, True, True), sb.dense_gemm(A, B, True, True))


# ---- eigen(S) and the polar factor ----------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [10, 64, 130, 500])
def test_eigh_matches_lapack(m):
    rng = np.random.default_rng(m)
    X = rng.standard_normal((3 * m, m))
    S = o.sample_cov(X)
    w, V, sweeps = sb.dense_eigh(S)
    wl = np.linalg.eigvalsh(S)[::-1]
    assert np.max(np.abs(w - wl)) < 1e-12 * wl[0]
    assert np.abs(V.T @ V - np.eye(m)).max() < 1e-13
    assert np.linalg.norm(S @ V - V * w[None, :]) < 1e-12 * wl[0] * np.sqrt(m)
    assert 1 <= sweeps <= 20


def test_eigh_spiked_spectrum_and_signs():
    X, Vq = o.spiked_samples(300, 700, 3, 0.1, seed=3)
    S = o.sample_cov(X)
    w, V, _ = sb.dense_eigh(S)
    wl, Vl = np.linalg.eigh(S)
    wl, Vl = wl[::-1], Vl[:, ::-1]
    assert np.max(np.abs(w - wl)) < 1e-12 * wl[0]
    assert np.all(colwise_inner(V[:, :3], Vl[:, :3]) > 1 - 1e-12)


def _svd_polar(B):
    u, s, vh = np.linalg.svd(B)
    return u @ vh, s


@pytest.mark.parametrize("m", [10, 100, 257])
def test_polar_well_conditioned(m):
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m, m)) + 2 * np.eye(m)
    U, sweeps = sb.dense_polar(B)
    Ur, s = _svd_polar(B)
    assert np.abs(U.T @ U - np.eye(m)).max() < 1e-13
    assert np.abs(U - Ur).max() < 50 * np.finfo(float).eps * (s[0] / s[-1]) * np.sqrt(m)
    H = U.T @ B
    assert np.abs(H - H.T).max() < 1e-12 * s[0]              # U^T B symmetric (positive semidefinite)
    assert 1 <= sweeps <= 20


def test_polar_graded_columns_like_the_procrustes_target():
    """Columns of wildly different scale (weights w0 * rho ~ 1e13 against O(1e3))."""
    rng = np.random.default_rng(7)
    m = 200
    B = rng.standard_normal((m, m))
    B[:, :3] *= 1e9
    U, _ = sb.dense_polar(B)
    Ur, s = _svd_polar(B)
    assert np.abs(U.T @ U - np.eye(m)).max() < 1e-13
    H = U.T @ B
    assert np.abs(H - H.T).max() < 1e-11 * s[0]
    assert np.linalg.eigvalsh((H + H.T) / 2).min() > -1e-10 * s[0]
    assert np.all(colwise_inner(U[:, :3], Ur[:, :3]) > 1 - 1e-12)


def test_polar_targets_of_a_real_run():
    m, n, q = 120, 300, 3
    X, _ = o.spiked_samples(m, n, q, 0.1, seed=3)
    S = o.sample_cov(X)
    Bs = []
    orig = oc.procrustes_target

    def spy(*a):
        B = orig(*a)
        Bs.append(B)
        return B
    oc.procrustes_target = spy
    try:
        oc.spEigenCov(S, q)
    finally:
        oc.procrustes_target = orig
    for B in Bs[::4] + [Bs[-1]]:
        U, _ = sb.dense_polar(B)
        Ur, s = _svd_polar(B)
        assert np.abs(U.T @ U - np.eye(m)).max() < 1e-13
        assert np.abs(U - Ur).max() < 200 * np.finfo(float).eps * (s[0] / s[-1])


# ---- spEigenCov end to end ---------------------------------------------------------------------------------------------
def _check_against_oracle(S, q, Vq=None, rho=0.5):
    r = sb.spEigenCov(S, q, rho)
    ro = oc.spEigenCov(S, q, rho, return_trace=True)
    F, Fo = r["info"]["F"], np.array(ro["trace"].F)
    assert len(F) == len(Fo), (len(F), len(Fo))
    assert np.array_equal(r["info"]["stage_of_k"], np.array(ro["trace"].stage))
    assert np.max(np.abs(F - Fo) / np.abs(Fo)) < 1e-9
    # the non-sparse eigenvalues inherit the conditioning of the m x m polar factor: measure the oracle's own noise floor
    # (the same computation on S perturbed by one ulp) next to the 1e-9 target
    rp = oc.spEigenCov(S * (1.0 + np.finfo(float).eps), q, rho)
    floor = np.max(np.abs(rp["values"] - ro["values"])) if len(rp["values"]) == len(ro["values"]) else 0.0
    assert np.max(np.abs(r["values"] - ro["values"])) < max(1e-9 * ro["values"].max(), 20 * floor)
    assert np.array_equal(r["vectors"][:, :q] != 0, ro["vectors"][:, :q] != 0)
    assert np.all(colwise_inner(r["vectors"][:, :q], ro["vectors"][:, :q]) > 1 - 1e-8)
    floor_cov = rel(rp["cov"], ro["cov"]) if len(rp["values"]) == len(ro["values"]) else 0.0
    assert rel(r["cov"], ro["cov"]) < max(1e-6, 20 * floor_cov)      # next to the oracle's own 1-ulp noise floor
    m = S.shape[0]
    assert np.abs(np.conj(r["vectors"].T) @ r["vectors"] - np.eye(m)).max() < 1e-8     # thresholding perturbs at ~thres
    if Vq is not None:
        assert np.all(colwise_inner(r["vectors"][:, :q], Vq) > 0.999)
    return r, ro


def test_speigencov_testthat_shape():
    """tests/testthat/test-sparseEigen.R:2-6,72-76: m = 10, n = 15, q = 3: output lengths m*m, m, m*m."""
    rng = np.random.default_rng(0)
    S = o.sample_cov(rng.standard_normal((15, 10)))
    r, _ = _check_against_oracle(S, 3)
    assert r["vectors"].size == 100 and r["values"].size == 10 and r["cov"].size == 100


def test_speigencov_m120():
    X, Vq = o.spiked_samples(120, 300, 3, 0.1, seed=3)
    _check_against_oracle(o.sample_cov(X), 3, Vq)


def test_speigencov_readme_example():
    """README.md:97-121 (m = 500, n = 600, q = 3)."""
    X, Vq = o.spiked_samples(500, 600, 3, 0.1, seed=42)
    S = o.sample_cov(X)
    r, _ = _check_against_oracle(S, 3, Vq)
    R = np.eye(500) + (Vq * (100.0 * np.arange(3, 0, -1) - 1.0)) @ Vq.T
    assert np.linalg.norm(r["cov"] - R) < 0.85 * np.linalg.norm(S - R)


def test_speigencov_config5_m2000():
    """BASELINE.json configs[4]: spEigenCov, m = 2000, n = 6000, q = 5 (the vignette shape scaled up)."""
    X, Vq = o.spiked_samples(2000, 6000, 5, 0.05, seed=42)
    S = sb.cov(X)
    assert rel(S, o.sample_cov(X)) < 1e-12
    r, ro = _check_against_oracle(S, 5, Vq)
    assert np.array_equal((r["vectors"][:, :5] != 0).sum(0), [100] * 5)
    assert r["info"]["gemms"] >= r["info"]["iterations"] and r["info"]["seconds_gemm"] > 0


def test_speigencov_q_variants():
    X, Vq = o.spiked_samples(200, 500, 8, 0.05, seed=11)
    S = o.sample_cov(X)
    _check_against_oracle(S, 8, rho=0.3)
    _check_against_oracle(S, 2, rho=0.9)
    # q = 1: the reference stops inside eigvalAlgo (undefined `tmp`, R/spEigenCov.R:312) whenever the first eigenvalue
    # needs no pooling; the drop-in pools {q} alone and completes
    with pytest.raises(oc.EigvalAlgoError):
        oc.spEigenCov(S, 1)
    r = sb.spEigenCov(S, 1)
    assert colwise_inner(r["vectors"][:, :1], Vq[:, :1])[0] > 0.999
    assert np.abs(r["vectors"].T @ r["vectors"] - np.eye(200)).max() < 1e-8


def test_speigencov_complex():
    """Complex Hermitian S (vignettes/SparseEigenvectors-vignette.Rmd:224-226): the real kernels on the embedded form."""
    for m, n, seed in [(10, 15, 0), (60, 150, 5), (130, 400, 9)]:
        X, Vq = o.spiked_samples(m, n, 3, 0.1 if m >= 30 else 0.3, seed=seed, cplx=True)
        S = o.sample_cov(X)
        r, ro = _check_against_oracle(S, 3)
        assert np.iscomplexobj(r["vectors"]) and np.iscomplexobj(r["cov"])
        assert np.abs(r["cov"] - np.conj(r["cov"].T)).max() < 1e-10 * np.abs(r["cov"]).max()


@pytest.mark.parametrize("name", ["real_m40", "real_m64_rho09", "cplx_m30"])
def test_speigencov_golden(name):
    """Committed fixtures (tests/golden/speigencov_cases.npz, tools/make_golden_cov.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "speigencov_cases.npz"))
    q = int(g[f"{name}.q"])
    r = sb.spEigenCov(g[f"{name}.S"], q, float(g[f"{name}.rho"]))
    F, Fg = r["info"]["F"], g[f"{name}.F"]
    assert len(F) == len(Fg) and np.max(np.abs(F - Fg) / np.abs(Fg)) < 1e-9
    assert np.array_equal(r["info"]["stage_of_k"], g[f"{name}.stage"])
    assert np.max(np.abs(r["values"] - g[f"{name}.values"])) < 1e-8 * g[f"{name}.values"].max()
    assert np.array_equal(r["vectors"][:, :q] != 0, g[f"{name}.vectors_q"] != 0)
    assert np.all(colwise_inner(r["vectors"][:, :q], g[f"{name}.vectors_q"]) > 1 - 1e-8)
    assert rel(r["cov"], g[f"{name}.cov"]) < 1e-6


def test_speigencov_errors():
    rng = np.random.default_rng(0)
    S = o.sample_cov(rng.standard_normal((15, 10)))
    Sna = S.copy(); Sna[1, 2] = Sna[2, 1] = np.nan
    for args in [(Sna, 3, 0.5), (S, float("nan"), 0.5), (S, 3, float("nan")), (S, -1, 0.5), (S, 0, 0.5), (S, 0.5, 0.5),
                 (S, 3, 0.0), (S + np.triu(np.ones((10, 10)), 1), 3, 0.5), (S, 11, 0.5),
                 (o.sample_cov(rng.standard_normal((5, 10))), 3, 0.5),            # low rank
                 (S - 2 * np.eye(10) * np.linalg.eigvalsh(S).max(), 3, 0.5)]:     # not PSD
        with pytest.raises(sb.SpEigenError):
            sb.spEigenCov(*args)
    with pytest.raises(sb.SpEigenError):
        sb.spEigenCov(np.cov(rng.standard_normal(15)), 3, 0.5)                    # univariate


# ---- spEigen(cov(X)) with cov(X) formed on the device -----------------------------------------------------------------
def test_speigen_cov_of_data_matches_host_cov():
    X, Vq = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    S = np.cov(X, rowvar=False)
    a = sb.spEigen(S, 3)
    b = sb.spEigen(X, 3, cov_of_data=True)
    k = min(len(a["info"]["F"]), len(b["info"]["F"]), 12)
    assert np.max(np.abs(a["info"]["F"][:k] - b["info"]["F"][:k]) / np.abs(a["info"]["F"][:k])) < 1e-10
    assert np.array_equal(a["vectors"] != 0, b["vectors"] != 0)
    assert np.all(colwise_inner(a["vectors"], b["vectors"]) > 1 - 1e-6)
    assert np.allclose(a["values"], b["values"], rtol=1e-11)


@pytest.mark.parametrize("world", [1, 2])
def test_solver_load_cov_of_data_shards(world):
    if sb.lib().speigen_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    rng = np.random.default_rng(3)
    n, m, q = 400, 700, 4
    X = rng.standard_normal((n, m))
    S = np.cov(X, rowvar=False)
    U = rng.standard_normal((m, q))
    with sb.Solver(m, q, world=world, local_gpus=world, n_gpus=world) as s:
        if world > 1:
            s.init_comm()
        s.load_cov_of_data(X)
        s.prepare()
        assert rel(s.contract(U), S @ U) < 1e-12
