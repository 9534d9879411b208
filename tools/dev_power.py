"""Sustained K1 loop with fine-grained clock/power sampling (development)."""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sparseeigen_b200 as sb
m = 50000
S = torch.randn(m, m, dtype=torch.float64, device="cuda")
s = sb.Solver(m, 32, check_symmetric=0, check_na=0, init_block=32)
s.init_comm(None); torch.cuda.synchronize(); s.load_device(S.data_ptr(), m); del S; torch.cuda.empty_cache()
def sample(tag, fn):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
    out = []
    t = threading.Thread(target=lambda: [out.append(l.strip()) for l in p.stdout], daemon=True); t.start()
    time.sleep(0.3)
    n0 = len(out)
    r = fn()
    n1 = len(out)
    time.sleep(0.1); p.terminate()
    rows = [l.split(", ") for l in out[n0:n1]]
    sm = np.array([float(x[0]) for x in rows]); pw = np.array([float(x[2]) for x in rows])
    print(f"{tag}: {r}; samples {len(rows)} sm MHz med {np.median(sm):.0f} min {sm.min():.0f} max {sm.max():.0f}; power W med {np.median(pw):.0f} max {pw.max():.0f}; temp {rows[-1][3]}; pcap {sum(x[4]=='Active' for x in rows)}/{len(rows)}", flush=True)
for ncols in (24, 20, 16):
    def f():
        ms = s.time_contract(ncols, 200, flush_l2=False)
        return f"ncols {ncols}: first10 {ms[:10].mean():.3f} last50 {ms[-50:].mean():.3f} ms"
    sample(f"ncols={ncols}", f)
    time.sleep(2)
