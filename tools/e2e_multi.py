"""End-to-end drop-in call on P GPUs of one process (what the R glue does): speigen_run_f64 on a host S (m = 50k, 20 GB),
from ordinary pageable memory (what R hands over) and from pinned memory, with the library's own time breakdown.

  GPUS=1,2,4,8 python tools/e2e_multi.py  >  profiles/r02_e2e_single_process_multi_gpu.jsonl
"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import sparseeigen_b200 as sb    # noqa: E402
from bench import host_cov, emit  # noqa: E402   (bench points the process stdout at stderr; emit() writes to the real one)

m, n, q = 50000, 10000, 20
KINDS = os.environ.get("KINDS", "pageable,pinned").split(",")
DEFER = os.environ.get("DEFER", "1").split(",")           # "1,0": A/B of the deferred host verdict (SPEIGEN_DEFER_VERIFY)
REPS = int(os.environ.get("REPS", "2"))
S_page = host_cov(m, n, q, 0.02, pinned=False).numpy()
S_pin = None
if "pinned" in KINDS:
    S_pin_t = torch.empty((m, m), dtype=torch.float64, pin_memory=True)
    S_pin_t.copy_(torch.from_numpy(S_page))
    S_pin = S_pin_t.numpy()
times = (ctypes.c_double * 6)()
for P in [int(x) for x in os.environ.get("GPUS", "1,2,4,8").split(",")]:
    sb.spEigen(np.asfortranarray(S_page[::25, ::25]), q, n_gpus=P)        # contexts + peer mappings of this GPU set (all 20 spikes kept)
    for kind in KINDS:
        S_host = S_page if kind == "pageable" else S_pin
        for defer in DEFER:
            os.environ["SPEIGEN_DEFER_VERIFY"] = defer
            per_rep = []
            for rep in range(REPS):
                t0 = time.perf_counter()
                r = sb.spEigen(S_host.T, q, n_gpus=P, check_symmetric=1)
                dt = time.perf_counter() - t0
                sb.lib().speigen_last_call_times(times)
                per_rep.append({"seconds": dt, "verdict_wait": times[0], "host_verify": times[1], "release_in_call": times[2], "in_library": times[3], "found_kept_solver": int(times[4])})
            i = r["info"]
            parts = {k: i[k] for k in ("seconds_setup", "seconds_upload", "seconds_prepare", "seconds_init", "seconds_loop")}
            emit({"n_gpus": P, "host_memory": kind, "deferred_host_verdict": defer == "1", "seconds_per_call": dt,
                  "mm_updates_per_s": 2 * i["iterations"] / dt, **parts,
                  "seconds_other": dt - sum(parts.values()), "reps": per_rep, "upload_staged": i["upload_staged"], "upload_half": i["upload_half"],
                  "iterations": i["iterations"], "polar_passes_max": i["polar_passes_max"], "F_final": float(i["F"][-1]),
                  "nonzeros": [int(x) for x in (r["vectors"] != 0).sum(axis=0)[:3]]})
