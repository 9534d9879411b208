"""The reference's own timing protocol (R_buildignore/package_comp_time.R:8-57): spEigen(X, q=3, rho=0.5, data=TRUE),
n = ceil(m/5), m = 100..800, mean wall time over repeated calls - here through the drop-in call with host buffers."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparseeigen_b200 as sb
from sparseeigen_b200 import synth
from oracle import speigen_oracle as o      # the CPU port timed beside the product (baseline leg only)
published = {100: 0.0140, 200: 0.0258, 300: 0.0342, 400: 0.0472, 500: 0.0614, 600: 0.0864, 700: 0.1206, 800: 0.1514}
rows = []
sb.spEigen(synth.spiked_samples(100, 20, 3, 0.1, seed=1)[0], 3, data=True)      # context warm-up
for m in range(100, 900, 100):
    n = -(-m // 5)
    X, Vq = synth.spiked_samples(m, n, 3, 0.1, seed=m)
    reps = 20
    ts = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        r = sb.spEigen(X, 3, rho=0.5, data=True)
        ts.append(time.perf_counter() - t0)
    first_call, ts = ts[0], ts[1:]          # the first call of a new shape re-creates the cached device context
    gpu = float(np.mean(ts))
    t0 = time.perf_counter()
    for _ in range(3):
        ro = o.spEigen(X, 3, rho=0.5, data=True)
    cpu = (time.perf_counter() - t0) / 3
    rec = np.abs(np.sum(r["vectors"] * Vq, axis=0)).min()
    same = bool(np.array_equal(r["vectors"] != 0, ro["vectors"] != 0))
    rows.append(dict(m=m, n=n, gpu_call_s=gpu, gpu_call_median_s=float(np.median(ts)), gpu_call_max_s=float(np.max(ts)), first_call_s=first_call, oracle_cpu_s=cpu, published_reference_s=published[m], iterations=r["info"]["iterations"],
                     loop_s=r["info"]["seconds_loop"], init_s=r["info"]["seconds_init"], upload_s=r["info"]["seconds_upload"],
                     min_recovery=float(rec), support_equal_oracle=same))
    print(json.dumps(rows[-1]), flush=True)
