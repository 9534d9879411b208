"""How far is the B200 path from the reference algorithm, and how far is the reference algorithm from ITSELF?

For every case: the CUDA path (through the C ABI), the oracle in both modes ("reference": G through the full eigenbasis Vx as
R/spEigen.R:177 does it; "direct": G = S V), and the same oracle on an input perturbed by one unit in the last place (the
reference algorithm's own noise floor).  Reported per pair: per-column |<u_a, u_b>| deficit, per-k relative deviation
of the objective trajectory, last iteration agreeing to 1e-10, final-F deviation, iteration / backtrack counts, support equality.

  python tools/parity_report.py > profiles/r02_parity.json          (needs a B200; the oracle legs run on the host)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np      # noqa: E402
import sparseeigen_b200 as sb    # noqa: E402
from oracle import speigen_oracle as o      # noqa: E402  (checker only)


def inner(A, B):
    return np.abs(np.sum(np.conj(A) * B, axis=0))


def perturb_1ulp(X, seed=0):
    rng = np.random.default_rng(seed)
    s = np.sign(rng.standard_normal(X.shape))
    Xp = X * (1 + np.finfo(float).eps * s)
    return Xp


def compare(a, b):
    """a, b: dict(vectors, F, a (step), bt, k)."""
    n = min(len(a["F"]), len(b["F"]))
    dev = np.abs(a["F"][:n] - b["F"][:n]) / np.maximum(1.0, np.abs(b["F"][:n]))
    agree = 0
    for x in dev:
        if x > 1e-10:
            break
        agree += 1
    first_big_a = next((i + 1 for i in range(n) if abs(a["a"][i]) > 1e3 or abs(b["a"][i]) > 1e3), None)
    return {"min_inner": float(inner(a["vectors"], b["vectors"]).min()),
            "inner_deficit": float(1.0 - inner(a["vectors"], b["vectors"]).min()),
            "support_identical": bool(np.array_equal(a["vectors"] != 0, b["vectors"] != 0)),
            "F_rel_dev_per_k": [float(f"{x:.3e}") for x in dev],
            "iterations_agreeing_to_1e-10": int(agree), "first_k_with_step_above_1e3": first_big_a,
            "F_final_rel_dev": float(abs(a["F"][-1] - b["F"][-1]) / max(1.0, abs(b["F"][-1]))),
            "iterations": [int(a["k"]), int(b["k"])], "backtracks": [int(a["bt"]), int(b["bt"])]}


def run_oracle(X, q, rho, data, mode):
    r = o.spEigen(X, q, rho=rho, data=data, mode=mode, return_trace=True)
    t = r["trace"]
    return {"vectors": r["vectors"], "F": np.array(t.F), "a": np.array(t.a_final), "bt": int(np.sum(t.backtracks)), "k": t.k,
            "values": r["values"]}


def run_gpu(X, q, rho, data):
    r = sb.spEigen(X, q, rho=rho, data=data)
    i = r["info"]
    return {"vectors": r["vectors"], "F": i["F"], "a": i["step_a"], "bt": int(i["backtracks"]), "k": int(i["iterations"]),
            "values": r["values"], "polar_passes_max": int(i["polar_passes_max"])}


def sym(S):
    return (S + np.conj(S.T)) / 2


CASES = []


def case(name):
    def deco(fn):
        CASES.append((name, fn))
        return fn
    return deco


@case("configs[0] README: cov m=500 n=100 q=3 real")
def _c0():
    X, _ = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    return o.sample_cov(X), 3, 0.5, False


@case("configs[1]: cov m=5000 n=1000 q=10 real")
def _c1():
    X, _ = o.spiked_samples(5000, 1000, 10, 0.05, seed=42)
    return o.sample_cov(X), 10, 0.5, False


@case("configs[2] scaled 1/50: data n=4000 x m=400 q=5 complex")
def _c2():
    X, _ = o.spiked_samples(400, 4000, 5, 0.05, seed=42, cplx=True)
    return X, 5, 0.5, True


@case("real data n=100 x m=500 q=3")
def _c3():
    X, _ = o.spiked_samples(500, 100, 3, 0.1, seed=42)
    return X, 3, 0.5, True


@case("complex cov m=400 q=3 rho=0.6")
def _c4():
    X, _ = o.spiked_samples(400, 500, 3, 0.2, seed=1, cplx=True)
    return o.sample_cov(X), 3, 0.6, False


@case("cov m=2000 n=400 q=10 real rho=0.3")
def _c5():
    X, _ = o.spiked_samples(2000, 400, 10, 0.05, seed=7)
    return o.sample_cov(X), 10, 0.3, False


def main():
    out = {"tolerances_north_star": {"inner": "1 - 1e-8", "objective_trajectory": 1e-10, "support": "identical"}, "cases": []}
    for name, fn in CASES:
        X, q, rho, data = fn()
        t0 = time.time()
        g = run_gpu(X, q, rho, data)
        od = run_oracle(X, q, rho, data, "direct")
        orf = run_oracle(X, q, rho, data, "reference")
        floors_d, floors_r = [], []
        for seed in range(3):                      # the floor itself is a coin flip (late accept/backtrack decisions): take 3 draws
            Xp = perturb_1ulp(X, seed)
            if not data:
                Xp = sym(Xp)
            floors_d.append(compare(run_oracle(Xp, q, rho, data, "direct"), od))
            floors_r.append(compare(run_oracle(Xp, q, rho, data, "reference"), orf))
        worst = lambda fl: max(fl, key=lambda c: c["inner_deficit"])
        rec = {"case": name, "gpu_polar_passes_max": g["polar_passes_max"],
               "values_rel_dev_gpu_vs_oracle": float(np.max(np.abs(g["values"] - od["values"]) / np.abs(od["values"]))),
               "gpu_vs_oracle_direct": compare(g, od), "gpu_vs_oracle_reference": compare(g, orf),
               "oracle_direct_vs_oracle_reference": compare(od, orf),
               "floor_oracle_direct_vs_1ulp": worst(floors_d), "floor_oracle_reference_vs_1ulp": worst(floors_r),
               "floor_inner_deficit_3_draws": {"direct": [c["inner_deficit"] for c in floors_d], "reference": [c["inner_deficit"] for c in floors_r]},
               "floor_F_final_rel_dev_3_draws": {"direct": [c["F_final_rel_dev"] for c in floors_d], "reference": [c["F_final_rel_dev"] for c in floors_r]},
               "seconds": round(time.time() - t0, 1)}
        out["cases"].append(rec)
        print(f"[parity] {name}: gpu-vs-direct deficit {rec['gpu_vs_oracle_direct']['inner_deficit']:.2e} (floor "
              f"{rec['floor_oracle_direct_vs_1ulp']['inner_deficit']:.2e}), gpu-vs-reference {rec['gpu_vs_oracle_reference']['inner_deficit']:.2e} "
              f"(floor {rec['floor_oracle_reference_vs_1ulp']['inner_deficit']:.2e}); F_final dev {rec['gpu_vs_oracle_direct']['F_final_rel_dev']:.2e} "
              f"(floor {rec['floor_oracle_direct_vs_1ulp']['F_final_rel_dev']:.2e}); agree-to-1e-10 k = "
              f"{rec['gpu_vs_oracle_direct']['iterations_agreeing_to_1e-10']} (floor {rec['floor_oracle_direct_vs_1ulp']['iterations_agreeing_to_1e-10']})",
              file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
