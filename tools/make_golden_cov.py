"""Generate tests/golden/speigencov_cases.npz from the oracle (oracle/speigencov_oracle.py).

The reference (R) cannot run in this image and ships no numeric golden vectors for spEigenCov, so these
fixtures pin the ORACLE against regressions; they are not outputs of the R package ("parity unpinned").
Run:  python tools/make_golden_cov.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import speigen_oracle as o
from oracle import speigencov_oracle as oc

out = {}


def add(name, S, q, rho=0.5):
    r = oc.spEigenCov(S, q, rho, return_trace=True, keep_iterates=True)
    t = r["trace"]
    out[f"{name}.S"] = S
    out[f"{name}.q"] = np.array(q)
    out[f"{name}.rho"] = np.array(rho)
    out[f"{name}.vectors_q"] = r["vectors"][:, :q]          # the sparse eigenvectors (the others are determined only up to
    out[f"{name}.values"] = r["values"]                       # the conditioning of the m x m polar factor)
    out[f"{name}.cov"] = r["cov"]
    out[f"{name}.F"] = np.array(t.F)
    out[f"{name}.stage"] = np.array(t.stage)
    out[f"{name}.rho_vec"] = r["rho_vec"]
    # eigvalAlgo known answers along the run: g = Re diag(-A^H S_hat A) is not stored by the trace, so re-derive one case
    vals = np.linalg.eigvalsh(S)[::-1]
    g = np.sort(np.random.default_rng(q).uniform(0.5, 2.0, S.shape[0]) / vals)
    g[:q] = g[:q][::-1]                                       # force the first q out of order -> pooling paths
    out[f"{name}.eigval_g"] = g
    out[f"{name}.eigval_x"] = np.array(vals[0])
    out[f"{name}.eigval_out"] = oc.eigval_algo(g, float(vals[0]), q)


X, _ = o.spiked_samples(40, 120, 3, 0.1, seed=21)
add("real_m40", o.sample_cov(X), 3)
X, _ = o.spiked_samples(64, 200, 2, 0.125, seed=22)
add("real_m64_rho09", o.sample_cov(X), 2, rho=0.9)
Xc, _ = o.spiked_samples(30, 100, 2, 0.2, seed=23, cplx=True)
add("cplx_m30", o.sample_cov(Xc), 2)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "speigencov_cases.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith(".F")})
