import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import sparseeigen_b200 as sb
from oracle import speigen_oracle as o, speigencov_oracle as oc
X, Vq = o.spiked_samples(200, 500, 8, 0.05, seed=11)
S = o.sample_cov(X)
Q, RHO = int(sys.argv[1]), float(sys.argv[2])
ro = oc.spEigenCov(S, Q, RHO, return_trace=True, keep_iterates=True)
for pol in (1, 0):
    r = sb.spEigenCov(S, Q, RHO, polish=pol)
    V, Vo = r["vectors"], ro["vectors"]
    print("polish", pol, "cov rel", np.linalg.norm(r["cov"] - ro["cov"]) / np.linalg.norm(ro["cov"]), "sweeps", r["info"]["svd_sweeps"],
          "F", np.max(np.abs(r["info"]["F"] - np.array(ro["trace"].F)) / np.abs(np.array(ro["trace"].F))),
          "vec inner min (all cols)", np.min(np.abs(np.sum(V * Vo, axis=0))), "orth", np.abs(V.T @ V - np.eye(200)).max(),
          "values rel", np.max(np.abs(r["values"] - ro["values"])) / ro["values"].max(), "k", len(r["info"]["F"]), len(ro["trace"].F))
    d = np.abs(np.sum(V * Vo, axis=0)); print("   worst columns", np.argsort(d)[:6], np.sort(d)[:6], "xi there", ro["values"][np.argsort(d)[:6]])
