"""A/B of the K1 variants in one process / one thermal state (development)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb
m = 50000
S = torch.randn(m, m, dtype=torch.float64, device="cuda")
s = sb.Solver(m, 32, check_symmetric=0, check_na=0, init_block=32)
s.init_comm(None); torch.cuda.synchronize(); s.load_device(S.data_ptr(), m); del S; torch.cuda.empty_cache()
ncols = int(os.environ.get("NCOLS", 20)); reps = int(os.environ.get("REPS", 150))
res = {"hybrid": [], "dmma": []}
s.time_contract(ncols, 100, flush_l2=False)   # heat up
for rnd in range(4):
    for name, env in (("hybrid", "1"), ("dmma", "0")):
        os.environ["SPEIGEN_HYBRID_DFMA"] = env
        ms = s.time_contract(ncols, reps, flush_l2=False)
        res[name].append(ms[20:].mean())
        print(f"round {rnd} {name}: mean {ms[20:].mean():.3f} ms  min {ms.min():.3f}  max {ms.max():.3f}", flush=True)
for k, v in res.items():
    print(k, "mean of rounds %.3f ms" % np.mean(v), "->", "%.0f GB/s" % ((8.0 * m * m + 16.0 * m * ncols) / np.mean(v) * 1e-6))
