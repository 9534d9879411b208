"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the oracle on the
same seeded inputs, against the golden fixtures, and through size-independent properties at full size.

Tolerances (FP64 path; north_star: |<u_ref, u_gpu>| >= 1 - 1e-8, identical support, objective trajectory within 1e-10):
  * single kernels <= 1e-13 relative, single MM update <= 1e-11, objective <= 1e-12 (golden known answers);
  * objective trajectory <= 1e-10 (north star) on EVERY iteration before the first SQUAREM step with |a| > 1e3 (k = 15..19
    of ~25: from there on U = V2 - 2 V1 + V is pure cancellation noise and a^2 U amplifies 1e-16 differences);
  * end to end, against the oracle in BOTH modes (G through the eigenbasis Vx as R/spEigen.R:177, and G = S V): identical
    support, |<u_ref, u_gpu>| >= 1 - 1e-7, final objective within 2e-6.  Measured (profiles/r02_parity.json, 6 cases x 2 modes):
    inner-product deficit 3e-11 .. 7.1e-8 (1e-8 met in 4 cases; the two others sit at 1.06e-8 and 7.1e-8), final objective
    1e-8 .. 6.7e-7 - the same range as the oracle against ITSELF on an input perturbed by one ulp (deficit 1e-15 .. 5.4e-8,
    final objective 8e-12 .. 5.3e-7), whose late accept/backtrack decisions flip just as the GPU's do.  The north-star 1e-8 /
    1e-10 end-to-end figures are below the reference algorithm's own reproducibility: which cases land under 1e-8 changes with
    any re-ordering of a floating-point sum (it did between two builds of this library), so the bound asserted here is 1e-7
    and the measured values are printed and recorded in profiles/r02_parity.json."""
import numpy as np
import pytest

import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
from conftest import colwise_inner

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def herm(rng, m, cplx):
    A = rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0)
    return (A + np.conj(A.T)) / 2


def panel(rng, m, q, cplx):
    return rng.standard_normal((m, q)) + (1j * rng.standard_normal((m, q)) if cplx else 0)


# ---------------------------------------------------------------------------------------------------
# K1: contraction
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,q", [(2, 1), (17, 3), (64, 8), (255, 9), (257, 16), (500, 3), (1000, 20), (2049, 1),
                                 (3000, 24), (1500, 33), (1024, 64)])
def test_contract_real_cov(m, q):
    rng = np.random.default_rng(m * 100 + q)
    S = herm(rng, m, False)
    with sb.Solver(m, min(q, m)) as s:
        s.load(S)
        U = panel(rng, m, min(q, m), False)
        assert rel(s.contract(U), S @ U) < 1e-13


@pytest.mark.parametrize("m,q", [(3, 1), (300, 3), (1001, 5), (777, 20), (512, 32)])
def test_contract_complex_cov(m, q):
    rng = np.random.default_rng(m + q)
    S = herm(rng, m, True)
    with sb.Solver(m, min(q, m), is_complex=True) as s:
        s.load(S)
        U = panel(rng, m, min(q, m), True)
        assert rel(s.contract(U), S @ U) < 1e-13


@pytest.mark.parametrize("n,m,q,cplx", [(15, 10, 3, False), (300, 200, 3, False), (1203, 515, 4, False), (64, 700, 8, False),
                                        (301, 200, 3, True), (2000, 333, 20, True)])
def test_contract_data(n, m, q, cplx):
    rng = np.random.default_rng(n + m)
    X = panel(rng, n, m, cplx)
    with sb.Solver(m, q, n=n, is_complex=cplx, is_data=True) as s:
        s.load(X)
        U = panel(rng, m, q, cplx)
        assert rel(s.contract(U), np.conj(X.T) @ (X @ U)) < 1e-13


def test_contract_hybrid_dfma_variant(monkeypatch):
    """The opt-in <NTF, RF=4> variant (left-over columns on dedicated DFMA warps) computes the same product."""
    rng = np.random.default_rng(9)
    for m, q in ((1300, 20), (700, 3), (900, 12)):
        S = herm(rng, m, False)
        U = panel(rng, m, q, False)
        with sb.Solver(m, q) as s:
            s.load(S)
            Y0 = s.contract(U)
            monkeypatch.setenv("SPEIGEN_HYBRID_DFMA", "1")
            Y1 = s.contract(U)
            monkeypatch.delenv("SPEIGEN_HYBRID_DFMA")
        assert rel(Y0, S @ U) < 1e-13 and rel(Y1, S @ U) < 1e-13


def test_contract_is_deterministic_and_linear():
    rng = np.random.default_rng(1)
    m, q = 4000, 20
    S = herm(rng, m, False)
    with sb.Solver(m, q) as s:
        s.load(S)
        U, W = panel(rng, m, q, False), panel(rng, m, q, False)
        Y1, Y2 = s.contract(U), s.contract(U)
        assert np.array_equal(Y1, Y2)                                  # fixed-order split-K fix-up
        assert rel(s.contract(2.0 * U - 3.0 * W), 2.0 * Y1 - 3.0 * s.contract(W)) < 1e-13


# ---------------------------------------------------------------------------------------------------
# K3: polar factor / small eigen-solver
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,q,cplx,cond", [(50, 3, False, 1e0), (1500, 10, False, 1e4), (3000, 20, False, 1e5),
                                           (700, 7, True, 1e3), (999, 20, True, 1e4), (2000, 64, False, 1e2)])
def test_polar(m, q, cplx, cond):
    rng = np.random.default_rng(q)
    A = panel(rng, m, q, cplx) * np.logspace(0, np.log10(cond), q)[None, :]
    with sb.Solver(m, q, is_complex=cplx) as s:
        U = s.polar(A)
    assert np.abs(np.conj(U.T) @ U - np.eye(q)).max() < 1e-13
    assert rel(U, o.polar(A)) < 1e-15 * cond * 50 + 1e-13


@pytest.mark.parametrize("m,q,cplx,cond", [(2000, 10, False, 1e9), (1500, 20, False, 1e12), (900, 6, True, 1e10)])
def test_polar_ill_conditioned(m, q, cplx, cond):
    """cond(B)^2 is beyond double precision: the Gram route needs more than two passes (eigenvalues drowned in round-off are
    clamped, every pass takes the square root of the condition number) and still returns THE polar factor - the reference's
    svd(G - H) -> u v^H (R/spEigen.R:180-182) always returns an orthonormal factor."""
    rng = np.random.default_rng(int(np.log10(cond)))
    Qm, _ = np.linalg.qr(panel(rng, m, q, cplx))
    Wm, _ = np.linalg.qr(panel(rng, q, q, cplx))
    A = (Qm * np.logspace(0, -np.log10(cond), q)[None, :]) @ np.conj(Wm.T)      # singular values 1 .. 1/cond
    with sb.Solver(m, q, is_complex=cplx) as s:
        U = s.polar(A)
        passes = s.tail_diag()["passes"]
    assert passes >= 3
    assert np.abs(np.conj(U.T) @ U - np.eye(q)).max() < 1e-13
    assert rel(U, Qm @ np.conj(Wm.T)) < 1e-15 * cond * 50 + 1e-13


def test_polar_rank_deficient():
    """A rank-deficient Procrustes target: R's svd() -> u v^H (R/spEigen.R:180-182) still returns an orthonormal factor (the
    completion of the null direction is arbitrary).  The Gram route does the same through its clamped passes: the factor is
    orthonormal and U^H A is Hermitian positive semi-definite (the defining property of a polar factor).  Only a target whose
    Gram matrix has an exactly zero row (a zero column) cannot be completed and is reported (SPEIGEN_ERR_NUMERIC)."""
    rng = np.random.default_rng(3)
    A = panel(rng, 500, 6, False)
    A[:, 4] = A[:, 1] - 2.0 * A[:, 2]          # exact linear dependence
    with sb.Solver(500, 6) as s:
        U = s.polar(A)
        assert s.tail_diag()["passes"] >= 3
        assert np.abs(U.T @ U - np.eye(6)).max() < 1e-12
        H = U.T @ A
        assert np.abs(H - H.T).max() < 1e-9 * np.abs(H).max() and np.linalg.eigvalsh((H + H.T) / 2).min() > -1e-9 * np.abs(H).max()
        assert rel(U @ H, A) < 1e-9
        Z = A.copy(); Z[:, 2] = 0.0
        with pytest.raises(sb.SpEigenError) as ei:
            s.polar(Z)
        assert ei.value.code == 26
        U = s.polar(panel(rng, 500, 6, False))  # the solver stays usable
        assert np.abs(U.T @ U - np.eye(6)).max() < 1e-13


@pytest.mark.parametrize("n,cplx", [(1, False), (2, False), (3, False), (8, False), (19, False), (24, False), (7, True), (20, True)])
def test_eig_small_jacobi(n, cplx):
    rng = np.random.default_rng(n)
    B = panel(rng, 3 * n + 2, n, cplx)
    Cm = np.conj(B.T) @ B
    with sb.Solver(64, min(n, 64), is_complex=cplx, init_block=n) as s:
        ev, W = s.eig_small(Cm)
    ref = np.linalg.eigvalsh(Cm)[::-1]
    assert np.allclose(ev, ref, rtol=1e-12, atol=1e-13 * ref.max())
    assert np.abs(Cm @ W - W * ev[None, :]).max() < 1e-12 * ref.max()
    assert np.abs(np.conj(W.T) @ W - np.eye(n)).max() < 1e-13


# ---------------------------------------------------------------------------------------------------
# single MM update / objective against the golden known answers and the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["real_cov", "real_data", "cplx_cov", "cplx_data"])
@pytest.mark.parametrize("fused", [True, False])
def test_mm_update_and_objective_golden(golden, name, fused):
    X, q = golden[f"{name}.X"], int(golden[f"{name}.q"])
    data = bool(golden[f"{name}.data"])
    cplx = np.iscomplexobj(X)
    prob = o.setup(X, q, float(golden[f"{name}.rho"]), data)
    pp, Eps, _ = o.schedule()
    n, m = X.shape
    with sb.Solver(m, q, n=n if data else 0, is_complex=cplx, is_data=data) as s:
        s.load(X)
        if data:
            s.prepare()                                   # centring happens on the device (R/spEigen.R:60)
        for st in (0, 5, 10):
            V1 = s.mm_update(prob.V, prob.d, prob.rho_vec, pp[st], Eps[st], fused=fused)
            assert rel(V1, golden[f"{name}.mm{st}"]) < 1e-11, (name, st)
            F = s.objective(golden[f"{name}.mm{st}"], prob.d, prob.rho_vec, pp[st], Eps[st])
            Fg = float(golden[f"{name}.obj{st}"])
            assert abs(F - Fg) <= 1e-12 * max(1.0, abs(Fg)), (name, st, F, Fg)


def test_fused_and_unfused_epilogue_agree_bitwise():
    X, _ = o.spiked_samples(900, 200, 5, 0.05, seed=4)
    S = o.sample_cov(X)
    prob = o.setup(S, 5, mode="direct")
    pp, Eps, _ = o.schedule()
    with sb.Solver(900, 5) as s:
        s.load(S)
        a = s.mm_update(prob.V, prob.d, prob.rho_vec, pp[3], Eps[3], fused=True)
        b = s.mm_update(prob.V, prob.d, prob.rho_vec, pp[3], Eps[3], fused=False)
    assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------------
# end to end against the oracle
# ---------------------------------------------------------------------------------------------------
def _e2e(X, q, data=False, rho=0.5, inner_tol=1e-7, f_tol=2e-6, **kw):
    """The drop-in call against the oracle in both modes.  inner_tol: bound on 1 - |<u_ref, u_gpu>| (see the module docstring)."""
    r = sb.spEigen(X, q, rho=rho, data=data, **kw)
    F = r["info"]["F"]
    G = np.conj(r["vectors"].T) @ r["vectors"]
    assert np.abs(np.diag(G) - 1).max() < 1e-12
    assert r["info"]["polar_passes_max"] == 2                  # no polar factor needed the ill-conditioned fallback
    out = None
    for mode in ("direct", "reference"):
        ro = o.spEigen(X, q, rho=rho, data=data, mode=mode, return_trace=True,
                       **{k: v for k, v in kw.items() if k in ("d", "V", "thres")})
        tr = ro["trace"]
        Fo = np.array(tr.F)
        # every iteration before the first huge SQUAREM step (in either run) must agree to the north-star 1e-10
        n = 0
        for k in range(min(len(F), len(Fo))):
            if abs(tr.a_final[k]) > 1e3 or abs(r["info"]["step_a"][k]) > 1e3:
                break
            n = k + 1
        assert n >= 12, (mode, n)
        dev = np.max(np.abs(F[:n] - Fo[:n]) / np.maximum(1.0, np.abs(Fo[:n])))
        assert dev < 1e-10, (mode, dev)
        assert np.array_equal(r["info"]["stage_of_k"][:n], np.array(tr.stage)[:n])
        assert np.array_equal(r["info"]["backtracks_per_k"][:n], np.array(tr.backtracks)[:n])
        assert np.array_equal(r["vectors"] != 0, ro["vectors"] != 0)                      # identical support
        deficit = 1 - colwise_inner(r["vectors"], ro["vectors"]).min()
        Fdev = abs(F[-1] - Fo[-1]) / abs(Fo[-1])
        print(f"e2e vs oracle[{mode}]: prefix {n} of {len(F)} iterations within {dev:.1e}; 1-|<u,u>| = {deficit:.2e}; final F {Fdev:.1e}")
        assert deficit <= inner_tol, (mode, deficit)
        assert Fdev <= f_tol, (mode, Fdev)
        assert np.allclose(r["values"], ro["values"], rtol=1e-11)
        assert np.all(colwise_inner(r["standard_vectors"], ro["standard_vectors"]) > 1 - 1e-10)
        if mode == "direct":
            out = ro
    return r, out


def test_e2e_readme_config():
    """configs[0]: README example spEigen(cov(X), q=3), m=500, n=100 (README.md:56-108)."""
    X, Vq = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    r, ro = _e2e(o.sample_cov(X), 3)
    assert np.all(colwise_inner(r["vectors"], Vq) >= 0.997)
    assert np.array_equal((r["vectors"] != 0).sum(axis=0), [50, 50, 50])
    # the oracle's own noise floor: same code, S perturbed by 1 ulp
    S = o.sample_cov(X)
    Sp = S * (1 + np.finfo(float).eps * np.sign(np.random.default_rng(0).standard_normal(S.shape)))
    Sp = (Sp + Sp.T) / 2
    rp = o.spEigen(Sp, 3, mode="direct")
    floor = 1 - colwise_inner(rp["vectors"], ro["vectors"]).min()
    ours = 1 - colwise_inner(r["vectors"], ro["vectors"]).min()
    print(f"noise floor (oracle vs 1-ulp-perturbed oracle): {floor:.2e}; GPU vs oracle: {ours:.2e}")
    assert ours < max(10 * floor, 1e-8)


def test_e2e_data_matrix():
    X, _ = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    _e2e(X, 3, data=True)


def test_e2e_config3_scaled_complex_data():
    """configs[2] at 1/50 scale: complex data matrix n = 4000 x m = 400, q = 5."""
    X, _ = o.spiked_samples(400, 4000, 5, 0.05, seed=42, cplx=True)
    _e2e(X, 5, data=True)


def test_e2e_config2_m5000_q10():
    """configs[1]: covariance input m=5000, q=10."""
    X, _ = o.spiked_samples(5000, 1000, 10, 0.05, seed=42)
    _e2e(o.sample_cov(X), 10)


def test_e2e_complex_cov_and_data():
    Xc, _ = o.spiked_samples(400, 500, 3, 0.2, seed=1, cplx=True)
    _e2e(o.sample_cov(Xc), 3, rho=0.6)
    _e2e(Xc, 3, data=True, rho=0.6)


def test_e2e_custom_d_V_thres_q1():
    X, _ = o.spiked_samples(300, 120, 4, 0.1, seed=8)
    S = o.sample_cov(X)
    prob = o.setup(S, 4)
    _e2e(S, 4, d=np.array([1.0, 0.9, 0.8, 0.7]))
    # a permuted, sign-flipped start takes 29 iterations, 10 of them in the round-off dominated regime: 1.5e-7 / 1.9e-6 measured
    _e2e(S, 4, V=prob.V[:, ::-1].copy() * np.array([1, -1, 1, -1])[None, :], inner_tol=1e-6, f_tol=1e-5)
    r = sb.spEigen(S, 1)
    assert r["vectors"].shape == (300, 1) and abs(np.linalg.norm(r["vectors"]) - 1) < 1e-12
    big = sb.spEigen(S, 4, thres=0.05)
    assert np.all((np.abs(big["vectors"]) == 0) | (np.abs(big["vectors"]) >= 0.05 * 0.99))
    # R/spEigen.R:160 `matrix(rep(nrm, m), ncol = q) * V` cycles the q factors through the column-major layout: when the
    # threshold removes real mass, entries are NOT scaled by their own column's factor.  Same iterate in (the solver is
    # bit-reproducible; thres = 0 returns the un-thresholded iterate), same quirk out.
    V = sb.spEigen(S, 4, thres=0.0)["vectors"]
    cut = sb.spEigen(S, 4, thres=0.17)["vectors"]
    want = o.finalize(V, 0.17)
    assert np.array_equal(cut != 0, want != 0) and 4 < np.count_nonzero(cut) < np.count_nonzero(V)
    assert np.max(np.abs(cut - want)) < 1e-12
    Vt = np.where(np.abs(V) < 0.17, 0.0, V)
    assert np.max(np.abs(want - Vt / np.linalg.norm(Vt, axis=0)[None, :])) > 1e-3     # the quirk is visible here


@pytest.mark.parametrize("name", ["real_cov", "real_data", "cplx_cov", "cplx_data"])
def test_e2e_golden(golden, name):
    X, q = golden[f"{name}.X"], int(golden[f"{name}.q"])
    r = sb.spEigen(X, q, rho=float(golden[f"{name}.rho"]), data=bool(golden[f"{name}.data"]))
    assert np.allclose(r["values"], golden[f"{name}.values"], rtol=1e-11)
    assert np.array_equal(r["vectors"] != 0, golden[f"{name}.vectors"] != 0)
    assert np.all(colwise_inner(r["vectors"], golden[f"{name}.vectors"]) > 1 - 1e-7)
    F, Fg = r["info"]["F"], golden[f"{name}.F"]
    assert np.allclose(F[:12], Fg[:12], rtol=1e-10)


def test_run_is_bit_reproducible():
    X, _ = o.spiked_samples(800, 200, 5, 0.05, seed=6)
    S = o.sample_cov(X)
    a, b = sb.spEigen(S, 5), sb.spEigen(S, 5)
    assert np.array_equal(a["vectors"], b["vectors"]) and np.array_equal(a["info"]["F"], b["info"]["F"])


def test_cached_context_carries_no_state_between_calls():
    """The one-shot call keeps a small solver alive between calls of the same shape; results must not depend on it."""
    Xa, _ = o.spiked_samples(600, 150, 4, 0.05, seed=31)
    Xb, _ = o.spiked_samples(600, 150, 4, 0.05, seed=32)
    Sa, Sb = o.sample_cov(Xa), o.sample_cov(Xb)
    sb.lib().speigen_shutdown()
    a1 = sb.spEigen(Sa, 4)                 # fresh context
    b1 = sb.spEigen(Sb, 4)                 # reused context, different data
    a2 = sb.spEigen(Sa, 4)                 # reused again
    low = Xa[:3].T @ Xa[:3]
    with pytest.raises(sb.SpEigenError, match="rank"): sb.spEigen(low, 4)     # error path on a reused context
    a3 = sb.spEigen(Sa, 4)
    sb.lib().speigen_shutdown()
    b2 = sb.spEigen(Sb, 4)                 # fresh context again
    for r in (a2, a3):
        assert np.array_equal(a1["vectors"], r["vectors"]) and np.array_equal(a1["info"]["F"], r["info"]["F"])
    assert np.array_equal(b1["vectors"], b2["vectors"]) and np.array_equal(b1["values"], b2["values"])


# ---------------------------------------------------------------------------------------------------
# the reference's behavioural tests at the boundary (tests/testthat/test-sparseEigen.R)
# ---------------------------------------------------------------------------------------------------
def test_reference_testthat_behaviours():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((15, 10))
    S = np.cov(X, rowvar=False)
    for args, kw in (((X, 3, 0.5), dict(data=True)), ((S, 3, 0.5), {})):               # :55-70 output lengths
        r = sb.spEigen(*args, **kw)
        assert r["vectors"].size == 30 and r["values"].size == 3 and r["standard_vectors"].size == 30
    # NA found by the device scan when the host pre-check is bypassed (call the C ABI directly)   :1-18
    import ctypes as C
    L = sb.lib()
    Sna = np.asfortranarray(S.copy()); Sna[2, 3] = np.nan; Sna[3, 2] = np.nan
    ob = sb._OutBuffers(10, 3, False)
    rc = L.speigen_run_f64(Sna.ctypes.data_as(sb._dp), 10, 10, 0, 3.0, 0.5, None, None, 1e-9, None, C.byref(ob.c))
    assert rc == 2
    A = S.copy(); A[0, 1] += 1e-3
    with pytest.raises(sb.SpEigenError, match="not symmetric"): sb.spEigen(A, 3)        # R/spEigen.R:75
    with pytest.raises(sb.SpEigenError, match="not PD"): sb.spEigen(S - 5 * np.eye(10), 3)   # :77
    low = X[:3].T @ X[:3]
    with pytest.raises(sb.SpEigenError, match="rank"): sb.spEigen(low, 5)               # :79
    with pytest.raises(sb.SpEigenError, match="rank"): sb.spEigen(X[:3], 5, data=True)  # :69


@pytest.mark.parametrize("decay", ["linear", "harmonic"])
def test_initialiser_on_slowly_separating_spectra(decay):
    """eigen()/svd() replacement (R/spEigen.R:68,76) on spectra without a spike: plain subspace iteration would need
    thousands of passes (lambda_25 / lambda_10 = 0.994 for the linear case); the Chebyshev-filtered iteration converges."""
    rng = np.random.default_rng(23)
    m, q = 1200, 10
    Qm, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = 1.0 - np.arange(m) / (2.0 * m) if decay == "linear" else 1.0 / (1.0 + 0.05 * np.arange(m))
    S = (Qm * lam[None, :]) @ Qm.T
    S = (S + S.T) / 2
    with sb.Solver(m, q, check_symmetric=1) as s:
        s.load(S)
        s.prepare()
        info = s.init()
        assert info["init_converged"] and info["init_passes"] < 200, info
        r = s.run(rho=0.5)
    ev = np.linalg.eigvalsh(S)[::-1][:q]
    assert np.allclose(r["values"], ev, rtol=1e-11)
    assert np.all(colwise_inner(r["standard_vectors"], Qm[:, :q]) > 1 - 1e-9)


def test_initialiser_when_q_reaches_into_the_bulk():
    """q larger than the number of separated eigenvalues: Vx[,1:q] (R/spEigen.R:76,86) then contains eigenvectors from the edge
    of a Marchenko-Pastur bulk whose relative gaps are ~1e-3 of the bulk width and ~1e-6 of the leading spike.  Plain or lightly
    filtered subspace iteration does not get there (round 1 carried on silently with 7e-4 wrong eigenvalues); the converged
    spikes are locked (deflated) and the rest is filtered with a high-degree Chebyshev polynomial."""
    rng = np.random.default_rng(11)
    m, n, q = 600, 3000, 12
    X = rng.standard_normal((n, m))
    for j, lam in enumerate((900.0, 500.0, 200.0)):          # three spikes on disjoint blocks of 30 variables
        v = np.zeros(m); v[j * 30:(j + 1) * 30] = 1 / np.sqrt(30)
        X += np.outer(X @ v, v) * (np.sqrt(lam) - 1.0)
    S = np.cov(X, rowvar=False)
    ev, U = np.linalg.eigh(S)
    ev, U = ev[::-1], U[:, ::-1]
    assert ev[2] > 50 * ev[3] and (ev[3] - ev[q]) / ev[3] < 0.2          # 3 spikes, then a tight cluster
    with sb.Solver(m, q, check_symmetric=1) as s:
        s.load(S)
        s.prepare()
        info = s.init()
        assert info["init_converged"] and info["init_passes"] < 100, info
        r = s.run(rho=0.5)
    assert np.allclose(r["values"], ev[:q], rtol=1e-11)
    assert np.all(colwise_inner(r["standard_vectors"], U[:, :q]) > 1 - 1e-9)
    # the MM loop from this start point follows the oracle's (started from LAPACK's eigenvectors) on the well-conditioned prefix;
    # the three spike columns end with the oracle's support (the nine bulk-edge columns are noise directions whose sparse
    # limits are not determined to a coin flip)
    ro = o.spEigen(S, q, mode="direct", return_trace=True)
    Fo = np.array(ro["trace"].F)
    print("F prefix deviation", np.abs(r["info"]["F"][:8] - Fo[:8]) / np.abs(Fo[:8]))
    assert np.allclose(r["info"]["F"][:6], Fo[:6], rtol=1e-8)
    assert np.array_equal(r["vectors"][:, :3] != 0, ro["vectors"][:, :3] != 0)


def test_not_pd_detected_beyond_the_top_block():
    """R/spEigen.R:77: a negative eigenvalue far below the leading block is found by the shifted power probe."""
    rng = np.random.default_rng(17)
    m = 400
    A = rng.standard_normal((m, 60))
    v = rng.standard_normal(m); v /= np.linalg.norm(v)
    S = A @ A.T / 60 - 2.0 * np.outer(v, v)
    S = (S + S.T) / 2
    assert np.linalg.eigvalsh(S)[0] < -1.0
    with pytest.raises(sb.SpEigenError, match="not PD"): sb.spEigen(S, 3)
    with pytest.raises(o.SpEigenError, match="not PD"): o.spEigen(S, 3)
    r = sb.spEigen(A @ A.T / 60, 3)                       # the PSD part alone passes (rank-deficient, lambda_min = 0)
    assert r["vectors"].shape == (m, 3)


def _flat_bottom(m, lam_min, rng):
    """PSD-looking matrix: 10 spikes, a flat bulk of eigenvalues in [0.01, 0.02], one eigenvalue lam_min."""
    Qm, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = np.concatenate([np.linspace(50, 5, 10), np.linspace(0.02, 0.01, m - 11), [lam_min]])
    S = (Qm * lam[None, :]) @ Qm.T
    return (S + S.T) / 2


@pytest.mark.parametrize("lam_min", [-25.0, -5.0])
def test_not_pd_clear_violations_are_found(lam_min):
    """R/spEigen.R:77 on the cases the Ritz-value probe can decide: |lambda_min| of the order of the leading eigenvalues."""
    S = _flat_bottom(300, lam_min, np.random.default_rng(5))
    with pytest.raises(o.SpEigenError, match="not PD"): o.spEigen(S, 3)
    with pytest.raises(sb.SpEigenError, match="not PD"): sb.spEigen(S, 3)


def test_not_pd_negative_diagonal_is_found_by_the_scan():
    """A diagonal entry below -1e-10 proves an eigenvalue below -1e-10 (x = e_i): found in the validation pass, whatever the spectrum."""
    X, _ = o.spiked_samples(300, 120, 3, 0.1, seed=2)
    S = o.sample_cov(X)
    S[17, :] *= 1e-4; S[:, 17] *= 1e-4                      # keeps S symmetric; tiny row/column
    S[17, 17] = -1e-6
    with pytest.raises(o.SpEigenError, match="not PD"): o.spEigen(S, 3)
    with pytest.raises(sb.SpEigenError, match="not PD"): sb.spEigen(S, 3)


@pytest.mark.xfail(strict=True, reason="documented limit: R/spEigen.R:77 tests the full spectrum (O(m^3) eigen()); this build bounds lambda_min "
                                       "with Ritz values from a few block power passes, which cannot resolve a -1e-6 eigenvalue sitting under a "
                                       "flat bulk (relative gap 1e-8: any polynomial method needs ~1e4 contractions); the matrix is accepted")
def test_not_pd_near_miss_under_a_flat_bulk():
    S = _flat_bottom(300, -1e-6, np.random.default_rng(5))
    with pytest.raises(o.SpEigenError, match="not PD"): o.spEigen(S, 3)
    with pytest.raises(sb.SpEigenError, match="not PD"): sb.spEigen(S, 3)


def test_unconverged_initialiser_is_an_error():
    """eigen()/svd() always converge; a subspace iteration that misses its tolerance must not be carried on silently."""
    X, _ = o.spiked_samples(600, 150, 4, 0.05, seed=3)
    S = o.sample_cov(X)
    with pytest.raises(sb.SpEigenError) as ei:
        sb.spEigen(S, 4, init_max_passes=2)
    assert ei.value.code == 27
    r = sb.spEigen(S, 4, init_max_passes=2, allow_unconverged_init=1)
    assert not r["info"]["init_converged"] and r["vectors"].shape == (600, 4)


def test_device_side_waits_are_bounded():
    """A peer that never publishes its exchange epoch (a rank that failed or left early) must not hang the GPU: the wait gives
    up after the budget, the call returns SPEIGEN_ERR_TIMEOUT and the solver object can still be destroyed / the library reused."""
    import time
    X, _ = o.spiked_samples(300, 90, 3, 0.1, seed=5)
    S = o.sample_cov(X)
    with sb.Solver(300, 3) as s:
        s.load(S)
        t0 = time.perf_counter()
        rc = s.L.speigen_solver_selftest_timeout(s.h, 200)
        dt = time.perf_counter() - t0
        assert rc == 28 and 0.15 < dt < 5.0, (rc, dt)
        assert b"waited more than" in s.L.speigen_last_error()
    r = sb.spEigen(S, 3)
    assert r["vectors"].shape == (300, 3)


# ---------------------------------------------------------------------------------------------------
# Hermitian-half upload (host_pack.h, Solver::load_host_half): one tile of every transposed pair crosses PCIe, the other is
# rebuilt on the device; isSymmetric / anyNA (R/spEigen.R:53,75) are decided on the host against the tiles that stayed there
class _env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        import os
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        import os
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _upload_info(s):
    import ctypes as C
    v = [C.c_int32() for _ in range(3)]
    assert s.L.speigen_solver_upload_info(s.h, *[C.byref(x) for x in v]) == 0
    return dict(staged=v[0].value, half=v[1].value, host_verified=v[2].value)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("no_staging", [0, 1])
def test_half_upload_rebuilds_the_matrix(cplx, no_staging):
    rng = np.random.default_rng(77)
    m, q = 1037, 6
    S = herm(rng, m, cplx)
    U = panel(rng, m, q, cplx)
    with _env(SPEIGEN_HALF_UPLOAD=0):
        with sb.Solver(m, q, is_complex=cplx) as s:
            s.load(S)
            assert _upload_info(s)["half"] == 0
            Y0 = s.contract(U)
    for block in (96, 100, 2048):        # ragged super-tiles, tiles off the 32-element grid, a single (diagonal) super-tile
        with _env(SPEIGEN_HALF_MIN_BYTES=1, SPEIGEN_HALF_BLOCK=block, SPEIGEN_NO_STAGING=no_staging):
            with sb.Solver(m, q, is_complex=cplx) as s:
                s.load(S)
                info = _upload_info(s)
                assert info["half"] == 1 and info["host_verified"] == 1 and info["staged"] == 1 - no_staging, info
                Y1 = s.contract(U)
        assert np.array_equal(Y0, Y1), (block, rel(Y1, Y0))
    assert rel(Y0, S @ U) < 1e-13


def test_half_upload_default_threshold_and_end_to_end():
    """Above 32 MB the half upload is the default path of the drop-in call: same bits as the full upload."""
    X, _ = o.spiked_samples(2500, 500, 4, 0.05, seed=31)
    S = o.sample_cov(X)
    S = (S + S.T) / 2                     # exactly symmetric, so that both uploads leave the same bits on the device
    with _env(SPEIGEN_HALF_UPLOAD=0):
        a = sb.spEigen(S, 4)
    with _env(SPEIGEN_HALF_BLOCK=512):
        b = sb.spEigen(S, 4)
    c = sb.spEigen(S, 4)
    assert not a["info"]["upload_half"] and b["info"]["upload_half"] and c["info"]["upload_half"]
    for r in (b, c):
        assert np.array_equal(a["vectors"], r["vectors"]) and np.array_equal(a["info"]["F"], r["info"]["F"])


@pytest.mark.parametrize("cplx", [False, True])
def test_half_upload_validation_matches_the_full_upload(cplx):
    """isSymmetric's tolerance (100 eps on sum|S - S^H| / sum|S|) and anyNA give the same verdict whichever side of the diagonal
    the offending entry sits on - also when its tile never leaves the host."""
    rng = np.random.default_rng(5)
    m = 400
    X = rng.standard_normal((3 * m, m)) + (1j * rng.standard_normal((3 * m, m)) if cplx else 0)
    S = np.conj(X.T) @ X / (3 * m)
    S = (S + np.conj(S.T)) / 2
    spots = [(3, 390), (390, 3), (150, 260), (260, 150), (200, 201)]
    for env in (dict(SPEIGEN_HALF_UPLOAD=0), dict(SPEIGEN_HALF_MIN_BYTES=1, SPEIGEN_HALF_BLOCK=64)):
        with _env(**env):
            for (i, j) in spots:
                A = S.copy(); A[i, j] += 1e-3
                with pytest.raises(sb.SpEigenError, match="not symmetric"): sb.spEigen(A, 3)
                A = S.copy(); A[i, j] += 1e-13          # below isSymmetric's tolerance: accepted, as in R
                assert sb.spEigen(A, 3)["vectors"].shape == (m, 3)
                import ctypes as C
                Sna = np.asfortranarray(S.copy()); Sna[i, j] = np.nan
                ob = sb._OutBuffers(m, 3, cplx)
                run = sb.lib().speigen_run_c64 if cplx else sb.lib().speigen_run_f64
                rc = run(Sna.ctypes.data_as(sb._dp), m, m, 0, 3.0, 0.5, None, None, 1e-9, None, C.byref(ob.c))
                assert rc == 2, (env, i, j, rc)


def test_one_shot_call_keeps_a_large_solver_alive_briefly():
    """Matrices above the cache threshold (256 MB; lowered to 0 here) keep their solver for SPEIGEN_KEEPALIVE_MS after a call: the
    next call of the same shape reuses it, an idle process gets the device memory back from the reaper thread.  Same bits always."""
    import ctypes as C
    import time
    X, _ = o.spiked_samples(900, 300, 3, 0.1, seed=3)
    S = o.sample_cov(X)
    t = (C.c_double * 6)()

    def call():
        r = sb.spEigen(S, 3)
        assert sb.lib().speigen_last_call_times(t) == 0
        return r, int(t[4]), int(t[5])
    assert sb.lib().speigen_shutdown() == 0
    with _env(SPEIGEN_CACHE_MAX_BYTES=0, SPEIGEN_KEEPALIVE_MS=0):
        a, kept, _ = call()
        b, kept_b, _ = call()
        assert kept == 0 and kept_b == 0                     # released inside the call
    with _env(SPEIGEN_CACHE_MAX_BYTES=0, SPEIGEN_KEEPALIVE_MS=500):
        c, kept_c, reaped0 = call()
        assert kept_c == 0
        time.sleep(2.0)                                      # idle: the reaper releases the kept solver
        assert sb.lib().speigen_last_call_times(t) == 0 and int(t[5]) == reaped0 + 1
    with _env(SPEIGEN_CACHE_MAX_BYTES=0, SPEIGEN_KEEPALIVE_MS=30000):
        d, kept_d, _ = call()
        e, kept_e, _ = call()
        assert kept_d == 0 and kept_e == 1                   # created, then found
        m = S.shape[0]
        t0 = time.perf_counter()
        with sb.Solver(m, 3) as s:                           # the resident API next to a kept one-shot solver
            s.load(S)
            assert rel(s.contract(np.eye(m, 3)), S[:, :3]) < 1e-13
        print("resident solver block: %.3f s" % (time.perf_counter() - t0))
        f, kept_f, _ = call()
        assert kept_f == 1
        assert sb.lib().speigen_shutdown() == 0              # stops the reaper, releases what is kept
    g, kept_g, _ = call()                                    # default thresholds: the small-matrix cache
    assert kept_g == 0
    for r in (b, c, d, e, f, g):
        assert np.array_equal(a["vectors"], r["vectors"]) and np.array_equal(a["info"]["F"], r["info"]["F"])


def test_max_iter_ends_a_stage_not_the_schedule():
    """R/spEigen.R:151: k >= max_iter only breaks the inner loop; the remaining (p, eps) stages still run two iterations each."""
    X, _ = o.spiked_samples(400, 100, 3, 0.1, seed=12)
    S = o.sample_cov(X)
    r = sb.spEigen(S, 3, max_iter=6)
    ro = o.spEigen(S, 3, mode="direct", return_trace=True, max_iter=6)
    assert r["info"]["iterations"] == ro["trace"].k
    assert np.array_equal(r["info"]["stage_of_k"], np.array(ro["trace"].stage))
    assert r["info"]["stage_of_k"][-1] == 11


# ---------------------------------------------------------------------------------------------------
# full-size properties (BASELINE config 4: m = 50000, q = 20) — no oracle at this size
# ---------------------------------------------------------------------------------------------------
def test_full_size_properties():
    import torch
    from bench import synth_cov_rows, symmetrize_
    m, n, q = 50000, 10000, 20
    S = symmetrize_(synth_cov_rows(m, n, q, 0.02, 0, m, torch.device("cuda", 0)))
    with sb.Solver(m, q) as s:
        s.init_comm(None)
        torch.cuda.synchronize()
        s.load_device(S.data_ptr(), m)
        rng = np.random.default_rng(0)
        U, W = rng.standard_normal((m, q)), rng.standard_normal((m, q))
        YU, YW = s.contract(U), s.contract(W)
        # linearity and symmetry of the operator: S(aU+bW) = aSU + bSW ;  W^T (S U) = (S W)^T U
        assert rel(s.contract(0.5 * U - 2.0 * W), 0.5 * YU - 2.0 * YW) < 1e-13
        assert np.abs(W.T @ YU - (U.T @ YW).T).max() / np.abs(W.T @ YU).max() < 1e-12
        # spot check 64 rows against a host dot product
        rows = rng.choice(m, 64, replace=False)
        Srows = S[torch.as_tensor(rows, device=S.device)].cpu().numpy()
        assert rel(YU[rows], Srows @ U) < 1e-13
        del S
        torch.cuda.empty_cache()
        s.prepare()
        info = s.init()
        assert info["init_converged"] and info["init_residual"] < 1e-12
        r = s.run(rho=0.5)
    V = r["vectors"]
    assert np.abs(V.T @ V - np.eye(q)).max() < 1e-6           # thresholded + renormalised orthonormal iterate
    assert np.array_equal((V != 0).sum(axis=0), np.full(q, 1000))     # planted cardinality 0.02 m
    for j in range(q):
        assert np.all(V[j * 1000:(j + 1) * 1000, j] != 0)
    F, st = r["info"]["F"], r["info"]["stage_of_k"]
    for k in range(1, len(F)):
        if st[k] == st[k - 1]:
            assert F[k] * (1 + np.sign(F[k]) * 1e-9) > F[k - 1]  # accepted steps are monotone within a stage
    assert np.all(np.diff(r["values"]) < 0)


# ---------------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs[2]: complex data matrix n = 200000 x m = 20000, q = 20, 64 GB) — no oracle at this size
# ---------------------------------------------------------------------------------------------------
def test_full_size_properties_config3_complex_data():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    n, m, q = 200000, 20000, 20
    if free < 1.15 * 16.0 * n * m:
        pytest.skip("needs ~75 GB of free device memory")
    card = 0.02
    sp = int(round(card * m))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    lam = 100.0 * torch.arange(q, 0, -1, dtype=torch.float64, device=dev)

    def cn(rows):   # CN(0,1) block [rows][n] in the library's layout (row = variable, contiguous = samples, interleaved complex)
        return torch.view_as_complex(torch.randn(rows, n, 2, dtype=torch.float64, device=dev, generator=g) / np.sqrt(2.0))

    with sb.Solver(m, q, n=n, is_complex=True, is_data=True) as s:
        s.init_comm(None)
        phases = torch.exp(1j * 2 * np.pi * torch.rand(q * sp, dtype=torch.float64, device=dev, generator=g)) / np.sqrt(sp)
        for c in range(q):            # planted block c: X += (Z v)(sqrt(lam_c) - 1) v^H   (vignette model, complex)
            Zb = cn(sp)
            v = phases[c * sp:(c + 1) * sp]
            zv = (Zb * v[:, None]).sum(dim=0)
            Zb += torch.conj(v)[:, None] * (zv * (torch.sqrt(lam[c]) - 1.0))[None, :]
            torch.cuda.synchronize()
            s.load_device_rows(c * sp, sp, Zb.data_ptr(), 2 * n)
            del Zb
        for r0 in range(q * sp, m, 1000):
            Zb = cn(min(1000, m - r0))
            torch.cuda.synchronize()
            s.load_device_rows(r0, Zb.shape[0], Zb.data_ptr(), 2 * n)
            del Zb
        torch.cuda.empty_cache()
        s.prepare()                                               # centring on the device (R/spEigen.R:60)
        rng = np.random.default_rng(0)
        U = rng.standard_normal((m, 4)) + 1j * rng.standard_normal((m, 4))
        W = rng.standard_normal((m, 4)) + 1j * rng.standard_normal((m, 4))
        YU, YW = s.contract(U), s.contract(W)                     # X^H (X .) : linear and Hermitian positive semi-definite
        assert rel(s.contract(0.5 * U - 2.0j * W), 0.5 * YU - 2.0j * YW) < 1e-13
        A = np.conj(W.T) @ YU
        assert np.abs(A - np.conj((np.conj(U.T) @ YW).T)).max() / np.abs(A).max() < 1e-12
        assert np.all(np.real(np.sum(np.conj(U) * YU, axis=0)) > 0)
        info = s.init()
        assert info["init_converged"] and info["init_residual"] < 1e-12
        r = s.run(rho=0.5)
    V = r["vectors"]
    assert np.abs(np.conj(V.T) @ V - np.eye(q)).max() < 1e-6
    assert np.array_equal((V != 0).sum(axis=0), np.full(q, sp))          # planted cardinality 0.02 m
    Vq = np.zeros((m, q), dtype=np.complex128)
    ph = phases.cpu().numpy()
    for c in range(q):
        Vq[c * sp:(c + 1) * sp, c] = ph[c * sp:(c + 1) * sp]
    assert np.all(colwise_inner(V, Vq) > 0.9999)
    F, st = r["info"]["F"], r["info"]["stage_of_k"]
    for k in range(1, len(F)):
        if st[k] == st[k - 1]:
            assert F[k] * (1 + np.sign(F[k]) * 1e-9) > F[k - 1]
    assert np.all(np.diff(r["values"]) < 0)
    assert r["info"]["polar_passes_max"] == 2
