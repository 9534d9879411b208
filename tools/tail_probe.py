"""Phase clock of the fused polar tail (tail.cu) and per-step times of the MM update on a resident spiked covariance.

  python tools/tail_probe.py M Q [STEPS] [P]   prints ms per MM update, the K1 share and CTA 0's phase clock of the last tail
                                               (P > 1: P GPUs driven by this one process, S row-sharded)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import sparseeigen_b200 as sb    # noqa: E402


def nvlink_kib(gpu=0):
    """Cumulative NVLink data counters of one GPU (KiB transmitted, received), summed over its links; None if unavailable."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    if not (tx or rx):
        print("nvidia-smi nvlink -gt d gave no counters:", out[:300].replace("\n", " | "), flush=True)
        return None
    return (tx, rx)


def spiked_cov(m, n, q, card, dev, seed=42):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    X = torch.randn(n, m, dtype=torch.float64, device=dev, generator=g)
    sp = max(1, int(round(card * m)))
    for j in range(q):
        blk = slice(j * sp, (j + 1) * sp)
        z = X[:, blk].sum(dim=1, keepdim=True) / np.sqrt(sp)
        X[:, blk] += z * ((np.sqrt(100.0 * (q - j)) - 1.0) / np.sqrt(sp))
    X -= X.mean(dim=0, keepdim=True)
    S = torch.empty(m, m, dtype=torch.float64, device=dev)
    for a in range(0, m, 8192):
        torch.matmul(X[:, a:a + 8192].T, X, out=S[a:a + 8192])
    S /= (n - 1)
    return S


if __name__ == "__main__":
    m, q = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    P = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda", 0)
    S = spiked_cov(m, max(200, m // 5), q, 0.02 if m >= 20000 else 0.05, dev)
    s = sb.Solver(m, q, world=P, local_gpus=P, check_symmetric=0)
    s.init_comm(None)
    torch.cuda.synchronize()
    for i in range(P):
        lo, hi = sb.shard_range(m, P, i)
        s.load_device(S[lo:hi].data_ptr(), m, local_idx=i)
    del S
    s.prepare()
    s.init()
    for stage in (0, 5, 10):
        s.mm_steps(5, stage=stage)
        s.kernel_timing(True)
        nv0 = nvlink_kib(0) if P > 1 else None
        ms = s.mm_steps(steps, stage=stage, timed=True)
        nv1 = nvlink_kib(0) if P > 1 else None
        k_ms, k_n = s.kernel_time()
        d = s.tail_diag()
        if nv0 and nv1:
            # per MM update on GPU 0 (the counters also see the stats-only tail + final copy of mm_steps: < 1 step's worth)
            print(f"NVLink GPU0 per MM update: tx {(nv1[0] - nv0[0]) * 1.024 / steps:.0f} kB, rx {(nv1[1] - nv0[1]) * 1.024 / steps:.0f} kB; "
                  f"algorithmic: rows to {P - 1} peers = {(P - 1) * (-(-m // P)) * q * 8 / 1e3:.0f} kB + mailbox words", flush=True)
        print(f"m={m} q={q} stage={stage}: {1e3 * ms / steps:.1f} us per MM update, K1 {1e3 * k_ms / max(k_n, 1):.1f} us, "
              f"rest {1e3 * (ms - k_ms) / steps:.1f} us; tail passes {d["passes"]} sweeps {d["sweeps"]} clock_us {d["clock_us"]} jacobi {d["jacobi_us"]}", flush=True)
    s.init()
    r = s.run(rho=0.5)
    i = r["info"]
    print(f"full solve: loop {1e3 * i['seconds_loop']:.2f} ms, {i['iterations']} outer iterations, {i['backtracks']} backtracks, "
          f"{i['contractions']} contractions -> {1e6 * i['seconds_loop'] / i['iterations']:.0f} us per outer iteration")
