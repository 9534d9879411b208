"""Times the block-Jacobi polar factor of a random 2048 x 2048 matrix (dev tool; tolerates non-convergence when the inner
sweeps are disabled through SPEIGEN_JACOBI_INNER=-1 to measure the fixed cost of a Jacobi round)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparseeigen_b200 as sb

rng = np.random.default_rng(0)
B = rng.standard_normal((2048, 2048))
for _ in range(2):
    t0 = time.perf_counter()
    try:
        U, sw = sb.dense_polar(B, jacobi_max_sweeps=int(os.environ.get("MAXSW", "40"))) if False else sb.dense_polar(B)
        print("sweeps", sw, "%.3f s" % (time.perf_counter() - t0))
    except Exception as e:
        print("err", str(e)[:80], "%.3f s" % (time.perf_counter() - t0))
