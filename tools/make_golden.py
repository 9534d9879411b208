"""Generate tests/golden/speigen_cases.npz from the oracle (oracle/speigen_oracle.py).

The reference (R) cannot run in this image and ships no numeric golden vectors, so these
fixtures pin the ORACLE against regressions; they are not outputs of the R package ("parity unpinned").
Run:  python tools/make_golden.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import speigen_oracle as o

out = {}
def add(name, X, q, data, rho=0.5, **kw):
    r = o.spEigen(X, q, rho=rho, data=data, mode="reference", return_trace=True, **kw)
    t = r["trace"]
    out[f"{name}.X"] = X
    out[f"{name}.q"] = np.array(q)
    out[f"{name}.rho"] = np.array(rho)
    out[f"{name}.data"] = np.array(int(data))
    out[f"{name}.vectors"] = r["vectors"]
    out[f"{name}.standard_vectors"] = r["standard_vectors"]
    out[f"{name}.values"] = r["values"]
    out[f"{name}.F"] = np.array(t.F)
    out[f"{name}.backtracks"] = np.array(t.backtracks)
    out[f"{name}.stage"] = np.array(t.stage)
    # single-step known answers at stages 1, 6, 11 from the start point
    prob = r["problem"]
    pp, Eps, _ = o.schedule()
    for st in (0, 5, 10):
        c1, c2, w0 = o.stage_constants(pp[st], Eps[st])
        V1 = o.mm_update(prob, prob.V, w0, c1, pp[st], Eps[st])
        out[f"{name}.mm{st}"] = V1
        out[f"{name}.obj{st}"] = np.array(o.objective(prob, V1, c1, c2, pp[st], Eps[st]))
    out[f"{name}.rho_vec"] = prob.rho_vec
    out[f"{name}.d"] = prob.d

X, _ = o.spiked_samples(60, 40, 3, 0.1, seed=11)
add("real_cov", o.sample_cov(X), 3, False)
add("real_data", X, 3, True)
Xc, _ = o.spiked_samples(40, 50, 2, 0.2, seed=12, cplx=True)
add("cplx_cov", o.sample_cov(Xc), 2, False)
add("cplx_data", Xc, 2, True, rho=0.6)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "speigen_cases.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
