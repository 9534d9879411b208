"""CPU tests (not gpu) for the spEigenCov row: the NumPy restatement of R/spEigenCov.R against the
reference's own behavioural tests and documented known-answers, and the host-side logic of the C ABI (validation,
schedule, eigvalAlgo) against the restatement.  No compute calls are made without a GPU."""
import numpy as np
import pytest

import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
from oracle import speigencov_oracle as oc


def _testthat_inputs(seed=0):
    """tests/testthat/test-sparseEigen.R:2-6: m = 10, n = 15, q = 3 (n > m so cov(X) is full rank)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((15, 10))
    return X, o.sample_cov(X)


# ---- tests/testthat/test-sparseEigen.R restated --------------------------------------------------------------------
def test_na_rejected():                                        # :9-17
    X, S = _testthat_inputs()
    Sna = S.copy(); Sna[1, 2] = np.nan; Sna[2, 1] = np.nan
    for args in [(Sna, 3, 0.5), (S, float("nan"), 0.5), (S, 3, float("nan"))]:
        with pytest.raises(o.SpEigenError):
            oc.spEigenCov(*args)


@pytest.mark.parametrize("q", [-1, 0, 0.5])
def test_q_must_be_positive_integer(q):                        # :33-35
    _, S = _testthat_inputs()
    with pytest.raises(o.SpEigenError):
        oc.spEigenCov(S, q, 0.5)
    assert sb.lib().speigencov_validate_args(10, 10, float(q), 0.5) == 3


def test_univariate_and_low_rank_rejected():                   # :47,:51
    X, _ = _testthat_inputs()
    with pytest.raises(o.SpEigenError):
        oc.spEigenCov(np.cov(X[:, 0]), 3, 0.5)                 # cov of one variable: 1 x 1
    rng = np.random.default_rng(1)
    Xs = rng.standard_normal((5, 10))                          # n < m: cov is rank deficient
    with pytest.raises(o.SpEigenError, match="low-rank"):
        oc.spEigenCov(o.sample_cov(Xs), 3, 0.5)
    assert sb.lib().speigencov_validate_args(1, 1, 3.0, 0.5) == 1


def test_output_lengths():                                     # :72-76
    _, S = _testthat_inputs()
    r = oc.spEigenCov(S, 3, 0.5)
    assert r["vectors"].shape == (10, 10) and r["values"].shape == (10,) and r["cov"].shape == (10, 10)


def test_validation_codes_follow_the_reference_order():
    L = sb.lib()
    assert L.speigencov_validate_args(10, 10, 11.0, 0.5) == 8            # q > m               R/spEigenCov.R:55
    assert L.speigencov_validate_args(10, 10, float("nan"), 0.5) == 2    # NA                  :56
    assert L.speigencov_validate_args(9, 10, 3.0, 0.5) == 6              # not square          :57
    assert L.speigencov_validate_args(10, 10, 3.0, 0.0) == 4             # rho <= 0            :59
    assert L.speigencov_validate_args(10, 10, 3.0, 0.5) == 0
    with pytest.raises(o.SpEigenError):
        oc.spEigenCov(np.eye(4), 5, 0.5)
    with pytest.raises(o.SpEigenError):
        oc.spEigenCov(np.eye(4) + np.triu(np.ones((4, 4)), 1), 2, 0.5)  # not symmetric
    with pytest.raises(o.SpEigenError):
        oc.spEigenCov(np.eye(4), 2, -1.0)


# ---- documented known-answers (statistical: R RNG stream not reproducible here) -----------------------------------
def test_readme_known_answer():
    """README.md:97-121: m = 500, n = 600, q = 3: recovery 0.9997197 0.9996029 0.9992848; covariance error 25.7 vs
    48.4 for the sample covariance."""
    m, n, q = 500, 600, 3
    X, Vq = o.spiked_samples(m, n, q, 0.1, seed=42)
    S = o.sample_cov(X)
    r = oc.spEigenCov(S, q, return_trace=True)
    rec = np.abs(np.sum(r["vectors"][:, :q] * Vq, axis=0))
    assert np.all(rec > 0.999)
    assert np.array_equal(r["vectors"][:, :q] != 0, Vq != 0)             # planted support exactly
    R = np.eye(m) + (Vq * (100.0 * np.arange(q, 0, -1) - 1.0)) @ Vq.T
    assert np.linalg.norm(r["cov"] - R) < 0.85 * np.linalg.norm(S - R)
    xi = r["values"]
    assert np.all(np.diff(xi[:q]) <= 0) and xi[q - 1] >= xi[q:].max() - 1e-12   # ordering constraints of the problem
    # objective decreases within every continuation stage (MM property)
    F, st = np.array(r["trace"].F), np.array(r["trace"].stage)
    for s in np.unique(st):
        f = F[st == s]
        assert np.all(np.diff(f) <= 1e-9 * np.abs(f[:-1]))
    # unitary iterate
    V = r["vectors"]
    assert np.abs(V.T @ V - np.eye(m)).max() < 1e-8


def test_vignette_known_answer():
    """vignettes/SparseEigenvectors-vignette.Rmd:148-186 (knitted html :201-219): m = 500, n = 600, q = 3, cardinality 0.2 m,
    rho = 0.6: recovery 0.9994329 0.9991827 0.9984716, covariance error 29.3 vs 48.4 for the sample covariance (R's RNG
    stream is not reproducible here, so the check is statistical: same model, three seeds)."""
    m, n, q = 500, 600, 3
    for seed in (42, 1, 2):
        X, Vq = o.spiked_samples(m, n, q, 0.2, seed=seed)
        S = o.sample_cov(X)
        r = oc.spEigenCov(S, q, 0.6)
        rec = np.abs(np.sum(r["vectors"][:, :q] * Vq, axis=0))
        assert np.all(rec > 0.998) and rec[0] > 0.9993
        assert np.array_equal((r["vectors"][:, :q] != 0).sum(0), [100, 100, 100])
        R = np.eye(m) + (Vq * (100.0 * np.arange(q, 0, -1) - 1.0)) @ Vq.T
        assert np.linalg.norm(r["cov"] - R) < 0.85 * np.linalg.norm(S - R)


# ---- host-side ABI logic against the restatement -------------------------------------------------------------------
def test_schedule_matches():
    p, e, t = sb.cov_schedule()
    po, eo, to = oc.schedule()
    assert np.allclose(p, po, rtol=1e-15) and np.allclose(e, eo, rtol=1e-15) and np.allclose(t, to, rtol=1e-15)
    cfg = sb.default_cov_cfg()
    assert (cfg.max_iter, cfg.n_continuation, cfg.p1, cfg.pk) == (3000, 5, 1.0, 5.0)    # R/spEigenCov.R:49,80-82
    assert (cfg.shift, cfg.rank_tol, cfg.psd_tol, cfg.rho_scale) == (1e-6, 1e-9, -1e-10, 10.0)


def test_eigval_algo_bit_exact_against_the_restatement():
    """eigvalAlgo (R/spEigenCov.R:164-338) in the C ABI vs the NumPy restatement: bit-identical on sorted, random, tied
    and partially violated inputs; both stop where R would (undefined `tmp`, non-terminating `while (1)`)."""
    rng = np.random.default_rng(0)
    old = oc.MAX_PASSES
    oc.MAX_PASSES = 100
    try:
        compared = errors = 0
        for trial in range(400):
            m = int(rng.integers(3, 30)); q = int(rng.integers(1, min(m, 8)))
            kind = trial % 4
            g = np.sort(rng.uniform(0.01, 5, m)) if kind == 0 else rng.uniform(0.01, 5, m)
            if kind == 2:
                g = np.round(g * 4) / 4 + 0.25
            if kind == 3:
                g = np.sort(rng.uniform(0.01, 5, m)); g[rng.integers(0, m, 3)] = rng.uniform(0.01, 5, 3)
            x = float(rng.uniform(1, 500))
            try:
                ref = oc.eigval_algo(g, x, q)
            except oc.EigvalAlgoError:
                errors += 1
                with pytest.raises(sb.SpEigenError):
                    sb.eigval_algo(g, x, q, strict_r=True)
                continue
            compared += 1
            assert np.array_equal(ref, sb.eigval_algo(g, x, q, strict_r=True))
        assert compared > 200 and errors > 10
    finally:
        oc.MAX_PASSES = old


def test_eigval_algo_on_the_trajectory_of_a_real_run():
    """Inputs eigvalAlgo actually sees: Re diag(-A^H S_hat A) along a spEigenCov run."""
    m, n, q = 60, 150, 3
    X, _ = o.spiked_samples(m, n, q, 0.1, seed=5)
    S = o.sample_cov(X)
    seen = []
    orig = oc.eigval_algo

    def spy(g, x, qq):
        out = orig(g, x, qq)
        seen.append((g.copy(), x, qq, out.copy()))
        return out
    oc.eigval_algo = spy
    try:
        oc.spEigenCov(S, q)
    finally:
        oc.eigval_algo = orig
    assert len(seen) >= 6
    for g, x, qq, out in seen:
        assert np.array_equal(out, sb.eigval_algo(g, x, qq))


def test_q1_pools_q_alone_where_r_would_stop():
    g = np.array([0.1, 0.5, 0.7, 0.9])
    with pytest.raises(oc.EigvalAlgoError):
        oc.eigval_algo(g, 10.0, 1)
    with pytest.raises(sb.SpEigenError):
        sb.eigval_algo(g, 10.0, 1, strict_r=True)
    out = sb.eigval_algo(g, 10.0, 1)
    phi = (1 + np.sqrt(1 + 40.0 * g)) / 20.0
    assert np.allclose(out, 1 / phi, rtol=1e-15)


def test_struct_layouts():
    import ctypes
    sz = (ctypes.c_int32 * 2)()
    assert sb.lib().speigencov_struct_sizes(ctypes.byref(sz, 0), ctypes.byref(sz, 4)) == 0
    assert tuple(sz) == (ctypes.sizeof(sb.CovCfg), ctypes.sizeof(sb.CovOut))


# ---- committed golden fixtures (tools/make_golden_cov.py; regression pins of the oracle, not R outputs) -------------------
GOLDEN_CASES = ["real_m40", "real_m64_rho09", "cplx_m30"]


def _golden_cov():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "speigencov_cases.npz"))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_the_golden_fixtures(name):
    g = _golden_cov()
    r = oc.spEigenCov(g[f"{name}.S"], int(g[f"{name}.q"]), float(g[f"{name}.rho"]), return_trace=True)
    q = int(g[f"{name}.q"])
    F = np.array(r["trace"].F)
    assert len(F) == len(g[f"{name}.F"]) and np.allclose(F, g[f"{name}.F"], rtol=1e-9)
    assert np.array_equal(np.array(r["trace"].stage), g[f"{name}.stage"])
    assert np.allclose(r["values"], g[f"{name}.values"], rtol=1e-8)
    assert np.array_equal(r["vectors"][:, :q] != 0, g[f"{name}.vectors_q"] != 0)
    inner = np.abs(np.sum(np.conj(r["vectors"][:, :q]) * g[f"{name}.vectors_q"], axis=0))
    assert np.all(inner > 1 - 1e-8)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_eigval_algo_golden(name):
    g = _golden_cov()
    out = sb.eigval_algo(g[f"{name}.eigval_g"], float(g[f"{name}.eigval_x"]), int(g[f"{name}.q"]))
    assert np.array_equal(out, g[f"{name}.eigval_out"])
