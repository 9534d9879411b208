import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb
P = int(os.environ.get("GPUS", 2)); n, m, q = 4000, 600, 4
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
Z = torch.view_as_complex(torch.randn(m, n, 2, dtype=torch.float64, device=dev, generator=g))   # [variable][sample]
Xh = Z.cpu().numpy().T.copy()            # n x m
rng = np.random.default_rng(0)
U = rng.standard_normal((m, q)) + 1j * rng.standard_normal((m, q))
ref = np.conj(Xh.T) @ (Xh @ U)
for mode in ("host", "device_rows"):
    s = sb.Solver(m, q, n=n, is_complex=True, is_data=True, world=P, local_gpus=P)
    s.init_comm(None)
    if mode == "host":
        s.load(Xh)
    else:
        torch.cuda.synchronize()
        for r in range(P):
            lo, hi = sb.shard_range(n, P, r)
            for r0 in range(0, m, 256):
                nr = min(256, m - r0)
                s.load_device_rows(r0, nr, Z.data_ptr() + 16 * (r0 * n + lo), 2 * n, local_idx=r)
    Y = s.contract(U)
    print(mode, "rel err", np.linalg.norm(Y - ref) / np.linalg.norm(ref), flush=True)
    s.close()
