"""The reference's quality protocol (R_buildignore/package_comp_param.R:8-58; vignette Figure "recovery"): m = 500, n = 200,
q = 3 sparse eigenvectors of cardinality 0.1 m, spEigen(X, q, rho, data = TRUE) for three values of rho, scored by
  angle recovery    |<u_j, v_j>|                                                    (:57)
  pattern recovery  1 - (1/m) sum_i ( 1{|u_ij| > 1e-6} - 1{|v_ij| > 1e-6} )          (:58, signed as in the reference)
through the drop-in call, with the oracle port beside it.  One JSON line per rho.
    python tools/recovery_vs_rho.py [--out FILE]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparseeigen_b200 as sb                      # noqa: E402
from sparseeigen_b200 import synth                 # noqa: E402
from oracle import speigen_oracle as o             # noqa: E402  (the CPU port scored beside the product: checker only)


def scores(U, Vq, m):
    angle = np.abs(np.sum(np.conj(U) * Vq, axis=0))
    pattern = 1.0 - ((np.abs(U) > 1e-6).astype(float) - (np.abs(Vq) > 1e-6).astype(float)).sum(axis=0) / m
    return angle.tolist(), pattern.tolist()


def run(m=500, n=200, rhos=(0.1, 0.5, 0.9)):
    q = 3
    X, Vq = synth.spiked_samples(m, n, q, 0.1, seed=42)
    lines = []
    for rho in rhos:
        r = sb.spEigen(X, q, rho, data=True)
        ro = o.spEigen(X, q, rho, data=True)
        ang, pat = scores(r["vectors"], Vq, m)
        ango, pato = scores(ro["vectors"], Vq, m)
        lines.append(dict(rho=rho, angle_recovery=ang, pattern_recovery=pat, oracle_angle_recovery=ango,
                          oracle_pattern_recovery=pato, support_equal_oracle=bool(np.array_equal(r["vectors"] != 0, ro["vectors"] != 0)),
                          standard_pca_angle=np.abs(np.sum(np.conj(r["standard_vectors"]) * Vq, axis=0)).tolist(),
                          iterations=int(r["info"]["iterations"])))
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--m", type=int, default=500)
    ap.add_argument("--n", type=int, default=200)
    a = ap.parse_args()
    lines = run(a.m, a.n)
    for ln in lines:
        print(json.dumps(ln), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
