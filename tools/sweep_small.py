"""The reference's own timing protocol (R_buildignore/package_comp_time.R:8-57): spEigen(X, q=3, rho=0.5, data=TRUE),
n = ceil(m/5), m = 100..800, mean wall time over repeated calls - here through the drop-in call with host buffers.
    python tools/sweep_small.py  >  profiles/r02_sweep_small_m100_800.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                 # noqa: E402
import sparseeigen_b200 as sb                      # noqa: E402
from sparseeigen_b200 import synth                 # noqa: E402
from oracle import speigen_oracle as o             # noqa: E402  (the CPU port timed beside the product: baseline leg only)

PUBLISHED = {100: 0.0140, 200: 0.0258, 300: 0.0342, 400: 0.0472, 500: 0.0614, 600: 0.0864, 700: 0.1206, 800: 0.1514}


def run(ms=range(100, 900, 100), reps=20, cpu_reps=3):
    rows = []
    sb.spEigen(synth.spiked_samples(100, 20, 3, 0.1, seed=1)[0], 3, data=True)      # context warm-up
    for m in ms:
        n = -(-m // 5)
        X, Vq = synth.spiked_samples(m, n, 3, 0.1, seed=m)
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            r = sb.spEigen(X, 3, rho=0.5, data=True)
            ts.append(time.perf_counter() - t0)
        first_call, ts = ts[0], ts[1:]          # the first call of a new shape re-creates the cached device context
        t0 = time.perf_counter()
        for _ in range(cpu_reps):
            ro = o.spEigen(X, 3, rho=0.5, data=True)
        cpu = (time.perf_counter() - t0) / cpu_reps
        rows.append(dict(m=m, n=n, gpu_call_s=float(np.mean(ts)), gpu_call_median_s=float(np.median(ts)), gpu_call_max_s=float(np.max(ts)),
                         first_call_s=first_call, oracle_cpu_s=cpu, published_reference_s=PUBLISHED.get(m), iterations=r["info"]["iterations"],
                         loop_s=r["info"]["seconds_loop"], init_s=r["info"]["seconds_init"], upload_s=r["info"]["seconds_upload"],
                         min_recovery=float(np.abs(np.sum(r["vectors"] * Vq, axis=0)).min()),
                         support_equal_oracle=bool(np.array_equal(r["vectors"] != 0, ro["vectors"] != 0))))
    return rows


if __name__ == "__main__":
    for row in run():
        print(json.dumps(row), flush=True)
