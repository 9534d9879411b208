"""Side-by-side trajectory trace: GPU path vs oracle (direct mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
X, Vq = o.spiked_samples(500, 100, 3, 0.1, seed=42)
S = o.sample_cov(X)
ro = o.spEigen(S, 3, mode="direct", return_trace=True); tr = ro["trace"]
for fused in (1, 0):
    print("=== fused", fused, flush=True)
    r = sb.spEigen(S, 3, fused_epilogue=fused, verbose=1)
    i = r["info"]
    print("k stage | a_gpu a_orc | bt_gpu bt_orc | F_gpu F_orc")
    for k in range(min(len(i["F"]), tr.k)):
        print(f"{k+1:2d} {i['stage_of_k'][k]:2d}/{tr.stage[k]:2d} | {i['step_a'][k]:12.4f} {tr.a_final[k]:12.4f} (init {tr.a_init[k]:12.4f}) | {i['backtracks_per_k'][k]} {tr.backtracks[k]} | {i['F'][k]:.15g} {tr.F[k]:.15g}  nR={tr.nR[k]:.3e} nU={tr.nU[k]:.3e}")
