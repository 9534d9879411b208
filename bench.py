#!/usr/bin/env python
"""bench.py — spEigen MM iterations/s and contraction-kernel roofline on B200.

Workload (BASELINE.json metric): spEigen covariance input m=50000, q=20, real FP64 (S = 20 GB),
row-sharded over N GPUs with an NCCL allgather of the m x q block per contraction.
A step = ONE MM update  V <- spEigenMMupdate(V)  (R/spEigen.R:167-183): one streamed contraction over S
with the reweighting fused in its epilogue + the polar (Procrustes) update.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: the oracle port, NumPy/OpenBLAS)

torch is used for plumbing only: device RNG/GEMM to synthesise S, torch.distributed for the bootstrap.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Libraries print to stdout behind our back (NCCL's "NCCL version ..." banner comes from a printf inside libnccl): the
# process-level stdout is pointed at stderr for the whole run and the ONE JSON line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

METRIC = "spEigen MM iters/sec (m=50k cov, q=20)"
UNIT = "MM iters/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.samples, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic_bytes(m, q, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the contraction kernel from the committed ncu --set full
    capture (profiles/r01_ncu_prof_contract_r01.csv), valid for the default single-GPU workload only; else None."""
    if (m, q, world) != (50000, 20, 1):
        return None
    p = os.path.join(ROOT, "profiles", "r01_ncu_prof_contract_r01.csv")
    try:
        rd = wr = None
        for line in open(p):
            f = line.strip().split(",")
            if f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            if f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        return rd + wr if rd is not None and wr is not None else None
    except OSError:
        return None


def synth_cov_rows(m, n, q, card, lo, hi, device, seed=42):
    """Rows lo..hi of the sample covariance of the spiked model of README.md:56-71, on `device`.
    Same seed on every rank -> every rank draws the same X."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.randn(n, m, dtype=torch.float64, device=device, generator=g)
    sp = int(round(card * m))
    lam = 100.0 * torch.arange(q, 0, -1, dtype=torch.float64, device=device)
    for j in range(q):   # X += (Z v_j)(sqrt(lam_j) - 1) v_j^T with v_j = 1/sqrt(sp) on its block
        blk = slice(j * sp, (j + 1) * sp)
        z = X[:, blk].sum(dim=1, keepdim=True) / np.sqrt(sp)
        X[:, blk] += z * ((torch.sqrt(lam[j]) - 1.0) / np.sqrt(sp))
    X -= X.mean(dim=0, keepdim=True)
    S = torch.empty(hi - lo, m, dtype=torch.float64, device=device)
    step = 8192
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        torch.matmul(X[:, a:b].T, X, out=S[a - lo:b - lo])
    S /= (n - 1)
    del X
    return S


def symmetrize_(S):
    import torch
    m = S.shape[0]
    step = 8192
    for a in range(0, m, step):
        for b in range(a, m, step):
            A = S[a:a + step, b:b + step]
            Bt = S[b:b + step, a:a + step].T
            avg = (A + Bt) * 0.5
            S[a:a + step, b:b + step] = avg
            S[b:b + step, a:a + step] = avg.T
    return S


def cpu_mm_update_time(S_host, q, steps, threads):
    """Oracle port (direct mode) timed on the host: MM updates/s on the same S."""
    from oracle import speigen_oracle as o
    m = S_host.shape[0]
    rng = np.random.default_rng(0)
    V = np.linalg.qr(rng.standard_normal((m, q)))[0]
    d = o.default_d(q)
    rho_vec = np.full(q, 0.5 * float(np.max(np.diag(S_host))))
    pp, Eps, _ = o.schedule()
    c1, c2, w0 = o.stage_constants(pp[0], Eps[0])
    prob = o.Problem(X=S_host, data=False, q=q, d=d, rho_vec=rho_vec, Vx=None, sv2=np.ones(q), V=V, nrow=m, mode="direct")
    V = o.mm_update(prob, V, w0, c1, pp[0], Eps[0])       # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        V = o.mm_update(prob, V, w0, c1, pp[0], Eps[0])
    return steps / (time.perf_counter() - t0)


def bench_speigencov(args):
    """Secondary workload (BASELINE.json configs[4]): spEigenCov, m = 2000, n = 6000, q = 5, real FP64.
    A step = one MM iteration of R/spEigenCov.R:97-124 (Procrustes update through the m x m polar factor + eigvalAlgo +
    objective).  Everything is measured through the drop-in C ABI call with HOST buffers (speigencov_run_f64), so `value`
    already contains the H2D of S and the D2H of vectors/values/cov; the roofline object is for the m x m x m FP64
    tensor-core GEMM (bound: FP64 tensor pipe, peak = DMMA rate measured by tools/microbench/fp64_pipes.cu on this pool)."""
    import sparseeigen_b200 as sb
    from sparseeigen_b200 import synth
    m, n, q = args.cov_m, args.cov_n, args.cov_q
    X, Vq = synth.spiked_samples(m, n, q, 0.05, seed=42)
    S = sb.cov(X)                                           # cov(X) on the device (symmetric DMMA GEMM)
    sb.spEigenCov(S[:256, :256].copy(), 2)                  # warm the context
    sampler = ClockSampler(0)
    sampler.start()
    calls, its, t_loop, t_gemm, gemms = max(1, args.steps // 10), 0, 0.0, 0.0, 0
    l0 = int(sb.lib().speigen_solver_launch_count(None))      # process-wide count of the library's kernel launches
    t0 = time.perf_counter()
    for _ in range(calls):
        r = sb.spEigenCov(S, q)
        its += r["info"]["iterations"]; t_loop += r["info"]["seconds_loop"]
        t_gemm += r["info"]["seconds_gemm"]; gemms += r["info"]["gemms"]
    dt = time.perf_counter() - t0
    launches = int(sb.lib().speigen_solver_launch_count(None)) - l0
    clocks = sampler.stop()
    mp = (m + 127) // 128 * 128
    tf = gemms * 2.0 * mp ** 3 / t_gemm / 1e12
    peak_tf = 37.0
    line = {"metric": "spEigenCov MM iters/sec (m=%d, n=%d, q=%d)" % (m, n, q), "value": its / dt, "unit": "MM iters/s", "n_gpus": 1,
            "steps": its, "warmup": 1, "ms_per_step": 1e3 * dt / its, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "spEigenCov covariance input m=%d (cov of n=%d samples), q=%d, real FP64, spiked model card=0.05 seed 42" % (m, n, q),
                       "step": "one MM iteration = 1 GEMM (S_hat V) + m x m polar factor (block Jacobi, 3 sweeps + 3 GEMMs) + eigvalAlgo + objective",
                       "l2": "m x m matrices (33.5 MB each) are L2-resident by design; the call re-uploads S and re-creates its context every time",
                       "timed_region": "whole speigencov_run_f64 calls incl. eigen(S), H2D of S, D2H of vectors/values/cov"},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": None,
                         "kernel": "k_gemm<1,1> (FP64 DMMA)", "ms_per_launch": 1e3 * t_gemm / gemms,
                         "algorithmic_flops_per_launch": 2.0 * mp ** 3, "launches_timed": gemms,
                         "peak_source": "FP64 tensor pipe measured on this pool (profiles/r01_microbench_fp64_pipes.log: DMMA.8x8x4 37.0 TFLOP/s); "
                                        "MEASURED_PEAKS.json holds no FP64 figure",
                         "note": "the GEMM is 4 % of the step; the step is bound by the latency of the Jacobi rounds (profiles/README.md)"},
            "e2e": {"value": its / dt, "unit": "MM iters/s", "h2d_bytes_per_step": int(8 * m * m * calls / its),
                    "d2h_bytes_per_step": int((2 * 8 * m * m + 8 * m) * calls / its), "calls": calls, "seconds_per_call": dt / calls,
                    "api": "speigencov_run_f64 (one-shot .Call entry) on a host S"},
            "gpu_launches": launches, "clocks": clocks,
            "full_solve": {"seconds_per_call": dt / calls, "seconds_loop": t_loop / calls, "iterations": its // calls,
                           "eig_sweeps": r["info"]["eig_sweeps"], "svd_sweeps": r["info"]["svd_sweeps"],
                           "recovery": [float(v) for v in np.abs(np.sum(r["vectors"][:, :q] * Vq, axis=0))]}}
    if not args.no_cpu:
        from oracle import speigencov_oracle as oc
        t0 = time.perf_counter()
        ro = oc.spEigenCov(S, q, return_trace=True)
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": len(ro["trace"].F) / tc, "unit": "MM iters/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "one full spEigenCov run of the oracle port (NumPy/OpenBLAS eigh + %d gesdd), %.1f s" % (len(ro["trace"].F), tc)}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=50000)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--q", type=int, default=20)
    ap.add_argument("--card", type=float, default=0.02)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-store exchange or NCCL allgather")
    ap.add_argument("--workload", default="speigen", choices=["speigen", "speigencov"],
                    help="speigen: the BASELINE metric (default); speigencov: BASELINE configs[4] (secondary line)")
    ap.add_argument("--cov-m", type=int, default=2000)
    ap.add_argument("--cov-n", type=int, default=6000)
    ap.add_argument("--cov-q", type=int, default=5)
    args = ap.parse_args()
    if args.workload == "speigencov" and args.impl == "ours":
        if int(os.environ.get("RANK", "0")) == 0:
            bench_speigencov(args)
        return
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    m, n, q = args.m, args.n, args.q
    workload = f"spEigen covariance input m={m}, q={q}, real FP64, spiked model n={n} card={args.card} seed 42"

    import torch
    # ---------------------------------------------------------------- reference arm (CPU) ------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        S = synth_cov_rows(m, n, q, args.card, 0, m, dev)
        S_host = S.cpu().numpy()
        del S
        steps = max(1, min(args.steps, 16))
        for _ in range(min(args.warmup, 1)):
            cpu_mm_update_time(S_host, q, 1, threads)
        t0 = time.perf_counter()
        v = cpu_mm_update_time(S_host, q, steps, threads)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": 1, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "note": "R is not installed in this image: the reference arm is the oracle port "
                           "(NumPy restatement of spEigenMMupdate, direct-S mode = 1 dgemm + svd per update; the R code does 2 dgemm "
                           "through Vx after an O(m^3) eigen() that cannot run at m=50k)", "blas": "OpenBLAS (numpy)"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{steps} MM updates on the full {m}x{m} S"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ---------------------------------------------------------------- our arm (B200) -----------------
    import sparseeigen_b200 as sb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=dev)
    lo, hi = sb.shard_range(m, world, rank)
    S = synth_cov_rows(m, n, q, args.card, lo, hi, dev)
    if world == 1:
        symmetrize_(S)
    solver = sb.Solver(m, q, world=world, rank=rank, local_gpus=1, first_device=local_rank,
                       check_symmetric=1 if world == 1 else 0, p2p_exchange=1 if args.exchange == "p2p" else 0)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sb.Solver.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        solver.init_comm(bytes(idt.cpu().numpy().tobytes()))
    else:
        solver.init_comm(None)
    torch.cuda.synchronize()                  # S was produced on torch's stream; the library copies on its own
    solver.load_device(S.data_ptr(), m)
    S_host = None
    want_host = (rank == 0 and world == 1 and (not args.no_e2e or not args.no_cpu))
    if want_host:
        S_host_t = torch.empty((m, m), dtype=torch.float64, pin_memory=True)
        S_host_t.copy_(S)
        S_host = S_host_t.numpy()
    del S
    torch.cuda.empty_cache()
    solver.prepare()
    init_info = solver.init()

    def barrier():
        solver.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    solver.mm_steps(args.warmup, stage=0)
    barrier()
    l0 = solver.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    solver.kernel_timing(True)
    barrier()
    # CUDA events on the library's launching stream bracket exactly K steps (synchronised on both sides)
    wall_ms = solver.mm_steps(args.steps, stage=0, timed=True)
    k_ms, k_n = solver.kernel_time()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launch_count() - l0
    # max over ranks (device-side work is bracketed by stream syncs; kernel times are CUDA events)
    tt = torch.tensor([wall_ms, k_ms / max(k_n, 1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms = float(tt[0]) / args.steps
    kern_ms = float(tt[1])
    value = 1e3 / step_ms
    peak, peak_src = peaks()
    rows_local = hi - lo
    alg_bytes = 8.0 * rows_local * m + 2 * 8.0 * m * q   # 8 m^2 / P + 16 m q per GPU
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    # real full solve on the resident matrix (reports the actual loop: outer iterations, contractions)
    solver.init()
    ts = time.perf_counter()
    res = solver.run(rho=0.5)
    solve_s = time.perf_counter() - ts
    info = res["info"]
    nz = (res["vectors"] != 0).sum(axis=0)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "sharding": f"S row-sharded over {world} GPU(s), exchange per contraction: {solver.exchange_mode()}",
                       "l2": "inputs larger than L2 (S shard = %.1f GB per GPU streamed every step)" % (8.0 * rows_local * m / 1e9),
                       "step": "one MM update = fused contraction+reweight, 2x(Gram, Jacobi, apply)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(m, q, world), "traffic_source": "profiles/r01_ncu_prof_contract_r01.csv (ncu --set full, one launch)",
                         "kernel": "k_contract_std<3,0>", "ms_per_launch": kern_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "launches_timed": k_n},
            "gpu_launches": int(launches),
            "full_solve": {"seconds_loop": info["seconds_loop"], "seconds_total": solve_s, "outer_iterations": info["iterations"],
                           "backtracks": info["backtracks"], "contractions": info["contractions"],
                           "outer_iters_per_s": info["iterations"] / info["seconds_loop"],
                           "init_passes": init_info["init_passes"], "init_seconds": init_info["seconds_init"],
                           "nonzeros_per_vector": [int(x) for x in nz[:4]], "F_final": float(info["F"][-1])},
            "clocks": clocks}
    if rank == 0 and world == 1:
        if not args.no_e2e:
            # end to end through the drop-in C ABI call with HOST buffers (H2D of S and D2H of results inside)
            solver.close()
            torch.cuda.empty_cache()
            calls = max(1, args.e2e_calls)
            sb.spEigen(np.asfortranarray(S_host[:2048, :2048]), q)   # warm the context on a small case
            t_e = time.perf_counter(); upd = 0
            for _ in range(calls):
                r = sb.spEigen(S_host.T, q, check_symmetric=1)   # .T of the C-ordered symmetric S = R's column-major view, no copy
                upd += 2 * r["info"]["iterations"]
            dt = time.perf_counter() - t_e
            line["e2e"] = {"value": upd / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * m * m),
                           "d2h_bytes_per_step": int(2 * 8 * m * q + 8 * q), "calls": calls, "seconds_per_call": dt / calls,
                           "mm_updates_per_call": upd // calls, "seconds_upload": r["info"]["seconds_upload"],
                           "seconds_init": r["info"]["seconds_init"], "seconds_loop": r["info"]["seconds_loop"],
                           "api": "speigen_run_f64 (one-shot .Call entry) on a pinned host S"}
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            cv = cpu_mm_update_time(S_host, q, args.cpu_steps, cores)
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_steps} MM updates (oracle port, direct-S mode, NumPy/OpenBLAS) on the full {m}x{m} S"}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
