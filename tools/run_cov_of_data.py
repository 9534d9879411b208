"""SURVEY §8(f)-2 at BASELINE configs[3] scale: spEigen(cov(X), q = 20) with X 10 000 x 50 000 on the host (4 GB).
cov(X) (column-centred, / (n - 1)) is formed on the device by the symmetric FP64 tensor-core GEMM straight into the
solver's resident shard (20 GB that never cross PCIe), then the usual prepare / init / MM loop run.
    python tools/run_cov_of_data.py [--m 50000 --n 10000 --q 20] [--out FILE]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparseeigen_b200 as sb                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=50000)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--q", type=int, default=20)
    ap.add_argument("--card", type=float, default=0.02)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from sparseeigen_b200 import synth
    t0 = time.time()
    X, Vq = synth.spiked_samples(a.m, a.n, a.q, a.card, seed=42)
    X = np.asfortranarray(X)
    t_gen = time.time() - t0
    # pure GEMM rate of the symmetric product (resident random operands, CUDA events)
    ms = sb.dense_gemm_time(a.m, a.m, a.n, True, True, symmetric=True, reps=3)
    flops_sym = (a.m / 128.0) * (a.m / 128.0 + 1) / 2 * 2.0 * 128 * 128 * a.n     # upper tiles only
    with sb.Solver(a.m, a.q) as s:
        t0 = time.time()
        s.load_cov_of_data(X)
        t_load = time.time() - t0
        t0 = time.time()
        s.prepare()
        t_prep = time.time() - t0
        ini = s.init()
        r = s.run(0.5)
    info = r["info"]
    rec = np.abs(np.sum(r["vectors"] * Vq, axis=0))
    # CPU reference point for the same product: a bounded dsyrk-like sample, extrapolated by flops
    k = min(a.m, 4000)
    Xs = np.ascontiguousarray(X[:, :k])
    t0 = time.time()
    _ = Xs.T @ Xs
    t_cpu = time.time() - t0
    cpu_tflops = 2.0 * k * k * a.n / t_cpu / 1e12
    line = dict(workload="spEigen(cov(X), q=%d): X %d x %d real FP64 on the host (spiked model card=%.2f seed 42)"
                         % (a.q, a.n, a.m, a.card),
                seconds_generate_host=t_gen, seconds_load_cov_of_data=t_load, seconds_prepare=t_prep,
                seconds_init=ini["seconds_init"], init_passes=ini["init_passes"], seconds_loop=info["seconds_loop"],
                outer_iterations=info["iterations"], backtracks=info["backtracks"], contractions=info["contractions"],
                gemm_ms=[float(v) for v in ms], gemm_tflops_symmetric=float(flops_sym / (min(ms) * 1e-3) / 1e12),
                h2d_bytes=int(X.nbytes), s_bytes_never_on_host=int(8 * a.m * a.m),
                recovery_min=float(rec.min()), nonzeros=[int(v) for v in (r["vectors"] != 0).sum(0)][:4],
                cpu_gemm_sample="X[:, :%d].T @ X[:, :%d] (NumPy/OpenBLAS, %d cores)" % (k, k, os.cpu_count()),
                cpu_gemm_tflops=cpu_tflops,
                cpu_cov_seconds_extrapolated=float(2.0 * a.m * a.m * a.n / 2 / (cpu_tflops * 1e12)))
    js = json.dumps(line)
    print(js)
    if a.out:
        with open(a.out, "w") as f:
            f.write(js + "\n")


if __name__ == "__main__":
    main()
