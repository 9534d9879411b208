"""K1 (streamed contraction) probes on a resident random 50k x 50k FP64 matrix, CUDA events per launch.

  python tools/k1_probe.py sweep            column sweep (COLS=8,16,20,24,32): HBM-bound vs FP64-pipe-bound regimes
  python tools/k1_probe.py ab               all-DMMA vs hybrid DMMA+DFMA variants in one process / one thermal state
  python tools/k1_probe.py power            200 back-to-back launches with nvidia-smi clock/power sampling
"""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import torch            # noqa: E402
import sparseeigen_b200 as sb    # noqa: E402


def resident(m):
    S = torch.randn(m, m, dtype=torch.float64, device="cuda")
    s = sb.Solver(m, 32, check_symmetric=0, check_na=0, init_block=32)
    s.init_comm(None)
    torch.cuda.synchronize()
    s.load_device(S.data_ptr(), m)
    del S
    torch.cuda.empty_cache()
    return s


def gbytes(m, ncols):
    return (8.0 * m * m + 16.0 * m * ncols) / 1e9


def sweep(s, m):
    for ncols in [int(x) for x in os.environ.get("COLS", "8,16,20,24,32").split(",")]:
        ms = s.time_contract(ncols, 24, flush_l2=False)
        print(f"ncols={ncols:2d}: first {ms[0]:.3f} ms, median {np.median(ms):.3f}, min {ms.min():.3f}, last8 {ms[-8:].mean():.3f} ms"
              f" -> {gbytes(m, ncols) / np.median(ms) * 1e3:.0f} GB/s, {2.0 * m * m * ncols / np.median(ms) * 1e-9:.2f} TFLOP/s", flush=True)


def ab(s, m):
    ncols, reps = int(os.environ.get("NCOLS", 20)), int(os.environ.get("REPS", 150))
    res = {"hybrid": [], "dmma": []}
    s.time_contract(ncols, 100, flush_l2=False)   # heat up
    for rnd in range(4):
        for name, env in (("hybrid", "1"), ("dmma", "0")):
            os.environ["SPEIGEN_HYBRID_DFMA"] = env
            ms = s.time_contract(ncols, reps, flush_l2=False)
            res[name].append(ms[20:].mean())
            print(f"round {rnd} {name}: mean {ms[20:].mean():.3f} ms  min {ms.min():.3f}  max {ms.max():.3f}", flush=True)
    for k, v in res.items():
        print(k, "mean of rounds %.3f ms -> %.0f GB/s" % (np.mean(v), gbytes(m, ncols) / np.mean(v) * 1e3))


def power(s, m):
    q = ("clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown")
    for ncols in (24, 20, 16):
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                             stdout=subprocess.PIPE, text=True)
        out = []
        threading.Thread(target=lambda: [out.append(l.strip()) for l in p.stdout], daemon=True).start()
        time.sleep(0.3)
        n0 = len(out)
        ms = s.time_contract(ncols, 200, flush_l2=False)
        n1 = len(out)
        time.sleep(0.1)
        p.terminate()
        rows = [l.split(", ") for l in out[n0:n1]]
        sm = np.array([float(x[0]) for x in rows])
        pw = np.array([float(x[2]) for x in rows])
        print(f"ncols={ncols}: first10 {ms[:10].mean():.3f} last50 {ms[-50:].mean():.3f} ms; sm MHz med {np.median(sm):.0f} min {sm.min():.0f}"
              f" max {sm.max():.0f}; power W med {np.median(pw):.0f} max {pw.max():.0f}; temp {rows[-1][3]};"
              f" sw_power_cap {sum(x[4] == 'Active' for x in rows)}/{len(rows)}", flush=True)
        time.sleep(2)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    m = int(os.environ.get("M", 50000))
    {"sweep": sweep, "ab": ab, "power": power}[mode](resident(m), m)
