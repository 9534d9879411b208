"""K1 timing sweep (development): per-launch CUDA-event times of back-to-back contractions at m=50k."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb
m = int(os.environ.get("M", 50000))
S = torch.randn(m, m, dtype=torch.float64, device="cuda")
s = sb.Solver(m, 32, check_symmetric=0, check_na=0, init_block=32)
s.init_comm(None); torch.cuda.synchronize(); s.load_device(S.data_ptr(), m); del S; torch.cuda.empty_cache()
for ncols in [int(x) for x in os.environ.get("COLS", "8,16,20,24,32").split(",")]:
    for flush in (False,):
        ms = s.time_contract(ncols, 24, flush_l2=flush)
        gb = (8.0 * m * m + 16.0 * m * ncols) / 1e9
        print(f"ncols={ncols:2d} flush={flush}: first {ms[0]:.3f} ms, median {np.median(ms):.3f} ms, min {ms.min():.3f}, last8 {ms[-8:].mean():.3f} ms -> {gb/np.median(ms)*1e3:.0f} GB/s; "
              f"flops {2.0*m*m*ncols/np.median(ms)*1e-9:.2f} TF", flush=True)
