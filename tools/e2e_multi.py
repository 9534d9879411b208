"""End-to-end drop-in call on P GPUs of one process (what the R glue does): speigen_run_f64 on a pinned host S."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb
from bench import synth_cov_rows, symmetrize_
m, n, q = 50000, 10000, 20
S = symmetrize_(synth_cov_rows(m, n, q, 0.02, 0, m, torch.device("cuda", 0)))
S_host_t = torch.empty((m, m), dtype=torch.float64, pin_memory=True); S_host_t.copy_(S); del S
torch.cuda.empty_cache()
S_host = S_host_t.numpy()
sb.spEigen(np.asfortranarray(S_host[:2048, :2048]), q)
for P in [int(x) for x in os.environ.get("GPUS", "1,2,4,8").split(",")]:
    for rep in range(2):
        t0 = time.perf_counter()
        r = sb.spEigen(S_host.T, q, n_gpus=P, check_symmetric=1)
        dt = time.perf_counter() - t0
    i = r["info"]
    print(json.dumps({"n_gpus": P, "seconds_per_call": dt, "mm_updates_per_s": 2 * i["iterations"] / dt, "upload_s": i["seconds_upload"],
                      "init_s": i["seconds_init"], "loop_s": i["seconds_loop"], "iterations": i["iterations"], "F_final": float(i["F"][-1]),
                      "nonzeros": [int(x) for x in (r["vectors"] != 0).sum(axis=0)[:3]]}), flush=True)
