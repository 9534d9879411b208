"""BASELINE configs[2]: spEigen data-matrix input X n=200000 x m=20000, q=20, complex FP64 (64 GB), on one B200.
Generates the spiked complex model on the device in variable blocks (library layout: row = variable, contiguous =
samples, interleaved complex), runs prepare / init / timed MM updates / a full solve, prints one JSON line."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb

n = int(os.environ.get("N", 200000)); m = int(os.environ.get("M", 20000)); q = int(os.environ.get("Q", 20))
card = 0.02; sp = int(round(card * m)); dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(42)
P = int(os.environ.get("GPUS", 1))
s = sb.Solver(m, q, n=n, is_complex=True, is_data=True, world=P, local_gpus=P)
s.init_comm(None)
shards = [sb.shard_range(n, P, r) for r in range(P)]
def load_rows(row0, nrows, Zb):
    torch.cuda.synchronize(0)                    # the library copies on its own streams
    for r, (lo, hi) in enumerate(shards):        # sample shard r keeps columns lo..hi of every variable row
        s.load_device_rows(row0, nrows, Zb.data_ptr() + 16 * lo, 2 * n, local_idx=r)
lam = 100.0 * torch.arange(q, 0, -1, dtype=torch.float64, device=dev)
def cn(rows):   # CN(0,1) block [rows][n]
    z = torch.randn(rows, n, 2, dtype=torch.float64, device=dev, generator=g) / np.sqrt(2.0)
    return torch.view_as_complex(z)
t0 = time.time()
# planted region: variables [0, q*sp)
phases = torch.exp(1j * 2 * np.pi * torch.rand(q * sp, dtype=torch.float64, device=dev, generator=g)) / np.sqrt(sp)
for c in range(q):
    Zb = cn(sp)                                  # [sp][n]
    v = phases[c * sp:(c + 1) * sp]              # entries of column c of Vq on its block
    zv = (Zb * v[:, None]).sum(dim=0)            # (Z Vq)[i, c] = sum_j Z[i,j] Vq[j,c]
    Zb += torch.conj(v)[:, None] * (zv * (torch.sqrt(lam[c]) - 1.0))[None, :]
    load_rows(c * sp, sp, Zb)
    del Zb
blk = 1000
for r0 in range(q * sp, m, blk):
    r = min(blk, m - r0)
    Zb = cn(r)
    load_rows(r0, r, Zb)
    del Zb
torch.cuda.synchronize(); torch.cuda.empty_cache()
t_gen = time.time() - t0
t0 = time.time(); s.prepare(); t_prep = time.time() - t0
info = s.init()
s.kernel_timing(True)
ms = s.mm_steps(3, stage=0, timed=True)         # warm-up
s.kernel_time()
K = int(os.environ.get("STEPS", 10))
ms = s.mm_steps(K, stage=0, timed=True)
k_ms, k_n = s.kernel_time()
s.init()
t0 = time.time(); res = s.run(rho=0.5); t_run = time.time() - t0
flops = 2 * 8.0 * n * m * q                      # two complex products per contraction
byts = 2 * 16.0 * n * m                          # X read twice (X*V then X^H*T)
Vq = np.zeros((m, q), dtype=np.complex128)
ph = phases.cpu().numpy()
for c in range(q):
    Vq[c * sp:(c + 1) * sp, c] = ph[c * sp:(c + 1) * sp]
rec = np.abs(np.sum(np.conj(res["vectors"]) * Vq, axis=0))
out = {"config": f"spEigen data-matrix input X n={n} x m={m}, q={q}, complex FP64 ({16.0*n*m/1e9:.0f} GB), {P} B200 (sample-sharded, single process)",
       "ms_per_mm_update": ms / K, "mm_updates_per_s": 1e3 * K / ms,
       "contraction_tflops": flops / (ms / K * 1e-3) / 1e12, "contraction_hbm_gbs_2pass": byts / (ms / K * 1e-3) / 1e9,
       "std_kernel_ms_per_launch": k_ms / max(k_n, 1), "init": info, "generate_s": t_gen, "prepare_s": t_prep,
       "full_solve_s": t_run, "outer_iterations": res["info"]["iterations"], "backtracks": res["info"]["backtracks"],
       "contractions": res["info"]["contractions"], "values_top3": [float(x) for x in res["values"][:3]],
       "nonzeros": [int(x) for x in (res["vectors"] != 0).sum(axis=0)[:5]],
       "recovery_min": float(rec.min()), "recovery_max": float(rec.max())}
print(json.dumps(out))
