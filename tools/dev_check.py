"""Quick GPU development check (not a test): kernels vs the oracle on small seeded inputs."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparseeigen_b200 as sb
from oracle import speigen_oracle as o

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
rng = np.random.default_rng(0)

def check_contract(m, q, cplx=False, data=False, n=0):
    if data:
        X = rng.standard_normal((n, m)) + (1j * rng.standard_normal((n, m)) if cplx else 0)
        s = sb.Solver(m, q, n=n, is_complex=cplx, is_data=True, check_symmetric=0)
        s.load(X)
        U = rng.standard_normal((m, q)) + (1j * rng.standard_normal((m, q)) if cplx else 0)
        Y = s.contract(U); ref = np.conj(X.T) @ (X @ U)
    else:
        A = rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0)
        S = (A + np.conj(A.T)) / 2
        s = sb.Solver(m, q, is_complex=cplx)
        s.load(S)
        U = rng.standard_normal((m, q)) + (1j * rng.standard_normal((m, q)) if cplx else 0)
        Y = s.contract(U); ref = S @ U
    print(f"contract m={m} q={q} cplx={cplx} data={data} n={n}: rel err {rel(Y, ref):.2e}", flush=True)
    s.close()

for (m, q) in [(64, 3), (500, 3), (1000, 20), (3000, 10), (777, 8), (2049, 1)]:
    check_contract(m, q)
check_contract(300, 3, cplx=True)
check_contract(1001, 5, cplx=True)
check_contract(200, 3, data=True, n=300)
check_contract(515, 4, data=True, n=1203)
check_contract(200, 3, cplx=True, data=True, n=301)

# polar + eig_small
m, q = 1500, 10
s = sb.Solver(m, q)
A = rng.standard_normal((m, q)) * np.logspace(0, 4, q)[None, :]
U = s.polar(A); print("polar rel", rel(U, o.polar(A)), "orth", np.abs(U.T @ U - np.eye(q)).max())
Cm = A.T @ A
ev, W = s.eig_small(Cm[:8, :8]); evr = np.linalg.eigvalsh(Cm[:8, :8])[::-1]
print("eig_small evals rel", np.abs(ev - evr).max() / evr.max(), "resid", np.abs(Cm[:8, :8] @ W - W * ev[None, :]).max() / evr.max())
s.close()

# one MM update + objective vs oracle
X, Vq = o.spiked_samples(500, 100, 3, 0.1, seed=42)
S = o.sample_cov(X)
prob = o.setup(S, 3, 0.5, mode="direct")
pp, Eps, tol = o.schedule()
s = sb.Solver(500, 3); s.load(S)
for st in (0, 5, 10):
    p, e = pp[st], Eps[st]
    c1, c2, w0 = o.stage_constants(p, e)
    V1o = o.mm_update(prob, prob.V, w0, c1, p, e)
    for fused in (True, False):
        V1 = s.mm_update(prob.V, prob.d, prob.rho_vec, p, e, fused=fused)
        print(f"mm_update stage {st} fused={fused}: rel {rel(V1, V1o):.2e}")
    Fo = o.objective(prob, V1o, c1, c2, p, e); Fg = s.objective(V1o, prob.d, prob.rho_vec, p, e)
    print(f"objective stage {st}: {Fo:.15g} vs {Fg:.15g} rel {abs(Fo-Fg)/abs(Fo):.2e}")
s.close()

# end to end
def e2e(S, q, data=False, **kw):
    t = time.time(); r = sb.spEigen(S, q, data=data, **kw); dt = time.time() - t
    ro = o.spEigen(S, q, data=data, mode="direct", return_trace=True)
    F = r["info"]["F"]; Fo = np.array(ro["trace"].F); n = min(len(F), len(Fo))
    inner = np.abs(np.sum(np.conj(r["vectors"]) * ro["vectors"], axis=0))
    print(f"e2e m={S.shape[1]} q={q} data={data}: {dt:.2f}s k={len(F)}/{len(Fo)} bt={r['info']['backtracks']}/{sum(ro['trace'].backtracks)} "
          f"init passes={r['info']['init_passes']} conv={r['info']['init_converged']} res={r['info']['init_residual']:.1e}")
    print("   F rel diff:", np.array2string(np.abs(F[:n] - Fo[:n]) / np.abs(Fo[:n]), precision=1, max_line_width=200))
    print("   inner:", inner, "support equal:", np.array_equal(r["vectors"] != 0, ro["vectors"] != 0),
          "values rel:", np.abs(r["values"] - ro["values"]).max() / ro["values"].max())
e2e(S, 3)
e2e(X, 3, data=True)
Xc, Vc = o.spiked_samples(200, 300, 3, 0.2, seed=1, cplx=True)
e2e(o.sample_cov(Xc), 3)
e2e(Xc, 3, data=True)
X2, _ = o.spiked_samples(2000, 400, 10, 0.05, seed=3)
e2e(o.sample_cov(X2), 10)
