#!/usr/bin/env python
"""bench.py — spEigen MM iterations/s and contraction-kernel roofline on B200.

Workload (BASELINE.json metric): spEigen covariance input m=50000, q=20, real FP64 (S = 20 GB), row-sharded over N GPUs.
A step = ONE MM update  V <- spEigenMMupdate(V)  (R/spEigen.R:167-183): one streamed contraction over S with the
reweighting fused in its epilogue (K1) + the fused polar tail kernel (Gram passes, q x q Jacobi/Taylor solve, applies,
column minima; row-sharded across the GPUs with peer-memory exchanges).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: the oracle port, NumPy/OpenBLAS on all host cores)

torch is used for plumbing only: device RNG/GEMM to synthesise S, torch.distributed for the bootstrap.
"""
from __future__ import annotations

import os
import sys

# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU legs of this script (reference arm, cpu_baseline) are
# meant to use every host core.  The BLAS thread pools read these variables when NumPy loads, so fix them first.
_HOST_CORES = os.cpu_count() or 1
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(_HOST_CORES)

import argparse          # noqa: E402
import hashlib           # noqa: E402
import json              # noqa: E402
import subprocess        # noqa: E402
import threading         # noqa: E402
import time              # noqa: E402

import numpy as np       # noqa: E402

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
DMMA_PEAK_TF = 37.0      # FP64 tensor pipe measured on this pool (profiles/r01_microbench_fp64_pipes.log)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def blas_threads():
    """Threads the BLAS behind NumPy will really use (threadpoolctl), next to the core count."""
    try:
        from threadpoolctl import threadpool_info
        info = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if info:
            return int(max(i["num_threads"] for i in info)), str(info[0].get("internal_api", "blas")) + " " + str(info[0].get("version", ""))
    except Exception:
        pass
    return _HOST_CORES, "OpenBLAS (numpy)"


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
    capture (profiles/r0?_ncu_prof_contract*.csv), valid for the default single-GPU workload only; else None."""
    if (m, q, world) != (50000, 20, 1):
        return None, None
    for name in ("r02_ncu_prof_contract.csv", "r01_ncu_prof_contract_r01.csv"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            rd = wr = None
            for line in open(p):
                f = line.strip().split(",")
                if f[0] == "dram__bytes_read.sum":
                    rd = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                if f[0] == "dram__bytes_write.sum":
                    wr = float(f[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            if rd is not None and wr is not None:
                return rd + wr, f"profiles/{name} (ncu --set full, one launch; not measured in this run)"
        except OSError:
            continue
    return None, None


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


def host_cov(m, n, q, card, pinned=False):
    """The full symmetric S on the host: synthesised on GPU 0 (identical bits to the sharded rows), copied down."""
    import torch
    dev = torch.device("cuda", 0)
    S = symmetrize_(synth_cov_rows(m, n, q, card, 0, m, dev))
    H = torch.empty((m, m), dtype=torch.float64, pin_memory=pinned)
    H.copy_(S)
    del S
    torch.cuda.empty_cache()
    return H


def cpu_mm_update_time(S_host, q, steps, warmup):
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
    for _ in range(max(warmup, 1)):
        V = o.mm_update(prob, V, w0, c1, pp[0], Eps[0])
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
    line = {"metric": "spEigenCov MM iters/sec (m=%d, n=%d, q=%d)" % (m, n, q), "value": its / dt, "unit": "MM iters/s", "n_gpus": 1,
            "steps": its, "warmup": 1, "ms_per_step": 1e3 * dt / its, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "spEigenCov covariance input m=%d (cov of n=%d samples), q=%d, real FP64, spiked model card=0.05 seed 42" % (m, n, q),
                       "step": "one MM iteration = 1 GEMM (S_hat V) + m x m polar factor (block Jacobi, 3 sweeps + 3 GEMMs) + eigvalAlgo + objective",
                       "l2": "m x m matrices (33.5 MB each) are L2-resident by design; the call re-uploads S and re-creates its context every time",
                       "timed_region": "whole speigencov_run_f64 calls incl. eigen(S), H2D of S, D2H of vectors/values/cov"},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": DMMA_PEAK_TF, "unit": "TFLOP/s", "frac": tf / DMMA_PEAK_TF, "traffic": None,
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
        line["cpu_baseline"] = {"value": len(ro["trace"].F) / tc, "unit": "MM iters/s", "cores": blas_threads()[0], "kind": "port",
                                "sample": "one full spEigenCov run of the oracle port (NumPy/OpenBLAS eigh + %d gesdd), %.1f s" % (len(ro["trace"].F), tc)}
    emit(line)


def e2e_calls(sb, S_host, q, n_gpus, calls, warm=1):
    """`calls` drop-in calls speigen_run_f64(S_host, ...) on n_gpus GPUs of THIS process after `warm` untimed ones (the first call
    of a process pins the staging ring and faults in the caller's pages); returns (seconds per call, last result)."""
    fallback = []

    def one():
        try:
            return sb.spEigen(S_host.T, q, n_gpus=n_gpus, check_symmetric=1)   # .T of the C-ordered symmetric S = R's column-major view, no copy
        except sb.SpEigenError as exc:       # never lose the line: retry once with the plain full-matrix upload and say so
            if os.environ.get("SPEIGEN_HALF_UPLOAD") == "0":
                raise
            fallback.append(str(exc))
            os.environ["SPEIGEN_HALF_UPLOAD"] = "0"
            return sb.spEigen(S_host.T, q, n_gpus=n_gpus, check_symmetric=1)
    first = None
    for _ in range(warm):
        t0 = time.perf_counter()
        one()
        if first is None:
            first = time.perf_counter() - t0
    import ctypes
    times = (ctypes.c_double * 6)()
    per_call = []
    t0 = time.perf_counter()
    r = None
    for _ in range(calls):
        tc = time.perf_counter()
        r = one()
        sec = time.perf_counter() - tc
        sb.lib().speigen_last_call_times(times)
        i = r["info"]
        per_call.append({"seconds": sec, "setup": i["seconds_setup"], "upload": i["seconds_upload"], "init": i["seconds_init"],
                         "loop": i["seconds_loop"], "host_verdict_wait": times[0], "host_verdict_work": times[1], "found_kept_solver": int(times[4])})
    dt = (time.perf_counter() - t0) / calls
    r["first_call_seconds"] = first
    r["upload_fallback"] = fallback[0] if fallback else None
    r["per_call"] = per_call
    return dt, r


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
    ap.add_argument("--no-qsweep", action="store_true")
    ap.add_argument("--e2e-calls", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-memory exchange or NCCL allgather")
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
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    m, n, q = args.m, args.n, args.q
    # one config for both arms (the driver compares them key by key)
    config = {"workload": f"spEigen covariance input m={m}, q={q}, real FP64, spiked model n={n} card={args.card} seed 42",
              "step": "one MM update V <- spEigenMMupdate(V) (R/spEigen.R:167-183) on the full m x m S",
              "l2": "inputs larger than L2 (the 20 GB S is streamed every step; no flush needed)"}

    import torch
    # ---------------------------------------------------------------- reference arm (CPU) ------------
    if args.impl == "reference":
        # Rank 0 alone runs it (the other ranks exit 0 without work).  R is not installed in this image and the reference is
        # pure R, so this arm times the oracle port: the NumPy restatement of spEigenMMupdate in direct-S mode (one dgemm + one
        # thin SVD per update) on every host core.  That is FLATTERING to the CPU: the R code does two dgemm through the full
        # eigenbasis Vx after an O(m^3) eigen() that cannot run at m = 50k on this host.
        if rank != 0:
            return
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        S = synth_cov_rows(m, n, q, args.card, 0, m, dev)
        S_host = S.cpu().numpy()
        del S
        threads, blas = blas_threads()
        v = cpu_mm_update_time(S_host, q, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "blas": blas, "host_cores": _HOST_CORES,
                                 "sample": f"{args.steps} MM updates (after {args.warmup} warm-up) of the oracle port on the full {m}x{m} S; "
                                           f"value does not depend on --gpus"},
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
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")      # host-side barriers: a waiting rank must not park a kernel on its GPU
    lo, hi = sb.shard_range(m, world, rank)
    S = synth_cov_rows(m, n, q, args.card, lo, hi, dev)
    if world == 1:
        symmetrize_(S)
    # isSymmetric.matrix (R/spEigen.R:75) runs in every mode: on one GPU locally, across ranks through the IPC-mapped shards
    solver = sb.Solver(m, q, world=world, rank=rank, local_gpus=1, first_device=local_rank,
                       check_symmetric=1, p2p_exchange=1 if args.exchange == "p2p" else 0)
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
    tail = solver.tail_diag()
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
    ncols_dmma = (q + 7) // 8 * 8
    flops_padded = 2.0 * rows_local * m * ncols_dmma

    # K1 alone for q in {8, 16, 20} (single GPU): the HBM-bound regime (q <= 16) next to the FP64-pipe-bound one (q = 20 -> 24)
    q_sweep = None
    if world == 1 and not args.no_qsweep:
        q_sweep = []
        for qq in (8, 16, 20):
            if qq > q:
                continue
            ms = np.median(solver.time_contract(qq, 12, flush_l2=False)[2:])
            ab = 8.0 * m * m + 16.0 * m * qq
            pad = (qq + 7) // 8 * 8
            q_sweep.append({"q": qq, "ms_per_launch": float(ms), "algorithmic_bytes_per_launch": ab, "achieved_gbs": ab / ms * 1e-6,
                            "frac_hbm": ab / ms * 1e-6 / peak, "dmma_columns": pad,
                            "frac_dmma_padded": 2.0 * m * m * pad / (ms * 1e-3) / 1e12 / DMMA_PEAK_TF})

    # real full solve on the resident matrix (reports the actual loop: outer iterations, contractions)
    solver.init()
    ts = time.perf_counter()
    res = solver.run(rho=0.5)
    solve_s = time.perf_counter() - ts
    info = res["info"]
    nz = (res["vectors"] != 0).sum(axis=0)

    traffic, traffic_src = ncu_traffic_bytes(m, q, world)
    config_ours = dict(config)
    config_ours["sharding"] = f"S row-sharded over {world} GPU(s); per MM update: K1 on the own rows, fused tail kernel on the own rows, " \
                              f"exchange: {solver.exchange_mode()} (q x q Gram partials + statistics through peer mailboxes, operand rows stored to every peer)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_ours,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "k_contract_std<3,0>", "ms_per_launch": kern_ms,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "launches_timed": k_n,
                         "second_ceiling": {"bound": "fp64 tensor pipe (q=20 padded to %d DMMA columns)" % ncols_dmma,
                                            "achieved_tflops": flops_padded / (kern_ms * 1e-3) / 1e12, "peak_tflops": DMMA_PEAK_TF,
                                            "frac": flops_padded / (kern_ms * 1e-3) / 1e12 / DMMA_PEAK_TF},
                         "q_sweep": q_sweep},
            "gpu_launches": int(launches),
            "tail_kernel": {"us_per_step_outside_K1": 1e3 * (step_ms - kern_ms), "gram_passes": tail["passes"], "jacobi_sweeps": tail["sweeps"],
                            "phase_clock_us_cta0": tail["clock_us"]},
            "full_solve": {"seconds_loop": info["seconds_loop"], "seconds_total": solve_s, "outer_iterations": info["iterations"],
                           "backtracks": info["backtracks"], "contractions": info["contractions"],
                           "outer_iters_per_s": info["iterations"] / info["seconds_loop"],
                           "init_passes": init_info["init_passes"], "init_seconds": init_info["seconds_init"],
                           "nonzeros_per_vector": [int(x) for x in nz[:4]], "F_final": float(info["F"][-1])},
            "clocks": clocks}

    # ---- multi-GPU self-check, part 1: every rank must hold the same result, bit for bit
    parity = None
    if world > 1:
        digest = hashlib.sha256(res["vectors"].tobytes() + info["F"].tobytes()).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, digest, group=cpu_group)
        parity = {"replicas_identical": len(set(digests)) == 1, "ranks": world}

    # ---- release the shards; the drop-in call below drives all N GPUs from rank 0's process
    solver.close()
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier(group=cpu_group)
    if rank == 0:
        need_host = (not args.no_e2e) or (world == 1 and not args.no_cpu)
        S_host = host_cov(m, n, q, args.card, pinned=False).numpy() if need_host else None
        if not args.no_e2e:
            # end to end through the drop-in C ABI call with HOST buffers: H2D of the 20 GB S, validation scan, initialiser, MM loop,
            # D2H of the results, on `world` GPUs of this process.  Headline: S in ordinary PAGEABLE memory (what R hands over).
            calls = max(1, args.e2e_calls)
            # warm the contexts / peer mappings of this GPU set on a small well-posed case: every 25th variable keeps all 20 spikes
            sb.spEigen(np.asfortranarray(S_host[::25, ::25]), q, n_gpus=world)
            dt, r = e2e_calls(sb, S_host, q, world, calls)
            upd = 2 * r["info"]["iterations"]
            h2d = int(sb.lib().speigen_last_upload_bytes())       # counted by the library from the copies it issued
            line["e2e"] = {"value": upd / dt, "unit": UNIT,
                           "h2d_bytes_per_step": int(h2d / upd), "d2h_bytes_per_step": int((2 * 8 * m * q + 8 * q) / upd),
                           "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": int(2 * 8 * m * q + 8 * q),
                           "host_bytes_read_per_call": int(8 * m * m),
                           "upload": ("Hermitian-half: one tile of every transposed 2048 x 2048 pair crosses PCIe, isSymmetric/anyNA "
                                      "are evaluated by the host workers against the tile that stays, the device rebuilds it"
                                      if r["info"]["upload_half"] else "full matrix"),
                           "calls": calls, "warmup_calls": 1, "first_call_seconds": r["first_call_seconds"],
                           "upload_fallback": r["upload_fallback"],
                           "seconds_per_call": dt, "per_call": r["per_call"], "n_gpus": world, "host_memory": "pageable",
                           "device_context": "the solver (shards, panels, streams, peer mappings) of a matrix above 256 MB is kept for 2 s after a "
                                             "call (SPEIGEN_KEEPALIVE_MS) and reused by a call of the same shape, then released by a reaper "
                                             "thread; `first_call_seconds` is a call that had to create it",
                           "mm_updates_counted_per_call": upd,
                           "counting": "2 MM updates (V1, V2) per outer iteration of R/spEigen.R:119-120; the call also runs "
                                       f"{r['info']['contractions']} streamed contractions in the loop (extrapolation/backtracking) and "
                                       f"{r['info']['init_passes']} initialiser passes, which are NOT counted",
                           "seconds_upload": r["info"]["seconds_upload"], "seconds_prepare": r["info"]["seconds_prepare"],
                           "seconds_init": r["info"]["seconds_init"], "seconds_loop": r["info"]["seconds_loop"],
                           "api": "speigen_run_f64 (one-shot .Call entry), single process driving %d GPU(s)" % world}
            # the same with a pinned host S (upper bound of the upload path)
            Hp = torch.empty((m, m), dtype=torch.float64, pin_memory=True)
            Hp.copy_(torch.from_numpy(S_host))
            dtp, rp = e2e_calls(sb, Hp.numpy(), q, world, 1)
            line["e2e"]["pinned_host"] = {"value": 2 * rp["info"]["iterations"] / dtp, "seconds_per_call": dtp,
                                          "seconds_upload": rp["info"]["seconds_upload"],
                                          "upload": "Hermitian-half" if rp["info"]["upload_half"] else "full matrix",
                                          "h2d_bytes_per_call": int(sb.lib().speigen_last_upload_bytes())}
            del Hp
            if world > 1:
                # ---- multi-GPU self-check, part 2: the N-GPU trajectory against the single-GPU one on the same S
                r1 = sb.spEigen(S_host.T, q, n_gpus=1, check_symmetric=1)
                F1, FN = r1["info"]["F"], info["F"]
                k = min(len(F1), len(FN), 10)
                inner = np.abs(np.sum(r1["vectors"] * res["vectors"], axis=0))
                parity.update({"F_prefix_len": int(k),
                               "F_prefix_max_rel_dev_vs_1gpu": float(np.max(np.abs(F1[:k] - FN[:k]) / np.abs(F1[:k]))),
                               "support_identical_vs_1gpu": bool(np.array_equal(r1["vectors"] != 0, res["vectors"] != 0)),
                               "min_abs_inner_vs_1gpu": float(inner.min()),
                               "iterations": [int(r1["info"]["iterations"]), int(info["iterations"])],
                               "ok": bool(parity["replicas_identical"] and np.array_equal(r1["vectors"] != 0, res["vectors"] != 0)
                                          and np.max(np.abs(F1[:k] - FN[:k]) / np.abs(F1[:k])) <= 1e-11 and inner.min() >= 1 - 1e-8)})
        if world == 1 and not args.no_cpu:
            cv = cpu_mm_update_time(S_host, q, args.cpu_steps, 1)
            threads, blas = blas_threads()
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port", "blas": blas, "host_cores": _HOST_CORES,
                                    "sample": f"{args.cpu_steps} MM updates (oracle port, direct-S mode, NumPy/OpenBLAS) on the full {m}x{m} S"}
        if parity is not None:
            line["parity_check"] = parity
        emit(line)
    if dist is not None:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
