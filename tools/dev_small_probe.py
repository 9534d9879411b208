import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
sb.spEigen(o.spiked_samples(100, 20, 3, 0.1, seed=1)[0], 3, data=True)
for m in (100, 200, 300, 800):
    n = -(-m // 5)
    X, Vq = o.spiked_samples(m, n, 3, 0.1, seed=m)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); r = sb.spEigen(X, 3, rho=0.5, data=True); ts.append(1e3 * (time.perf_counter() - t0))
    print(m, ["%.1f" % t for t in ts], "loop %.2f init %.2f up %.2f" % (1e3*r["info"]["seconds_loop"], 1e3*r["info"]["seconds_init"], 1e3*r["info"]["seconds_upload"]), "passes", r["info"]["init_passes"], flush=True)
