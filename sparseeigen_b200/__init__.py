"""sparseeigen_b200 — host-side mirror of sparseEigen::spEigen() over the B200 CUDA library.

The reference's only interface for this path is the R closure
    spEigen(X, q = 1, rho = 0.5, data = FALSE, d = NA, V = NA, thres = 1e-9)     (R/spEigen.R:46)
returning list(vectors, standard_vectors, values) (R/spEigen.R:162).  R is not available in the build
image, so this module plays the role of the R wrapper: same argument names, defaults, meaning and error
conditions, forwarding to the C ABI in include/speigen_b200.h (libspeigen_b200.so) through ctypes.
There is NO CPU fallback: if the CUDA library is missing or no B200 is visible the call raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np

__all__ = ["spEigen", "spEigenCov", "SpEigenError", "Solver", "lib", "library_path", "default_cfg", "schedule", "default_d",
           "shard_range", "panel_pitch", "default_cov_cfg", "cov_schedule", "eigval_algo", "dense_gemm", "dense_gemm_time",
           "dense_polar", "dense_eigh", "cov"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libspeigen_b200.so"


def library_path() -> str:
    # SPEIGEN_LIB: an alternative build of the same library (kernel A/B experiments); default: the in-tree build
    return os.environ.get("SPEIGEN_LIB") or os.path.join(_HERE, _LIB_NAME)


class SpEigenError(ValueError):
    """Raised where the reference calls stop() (R/spEigen.R:52-55,69,75,77,79) and for device errors."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Cfg(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32), ("n_continuation", C.c_int32), ("p1", C.c_double), ("pk", C.c_double),
        ("tol_factor", C.c_double), ("backtrack_slack", C.c_double), ("rank_tol", C.c_double), ("pd_tol", C.c_double),
        ("max_backtracks", C.c_int32), ("init_max_passes", C.c_int32), ("init_tol", C.c_double),
        ("init_block", C.c_int32), ("init_seed", C.c_uint64), ("check_symmetric", C.c_int32), ("check_na", C.c_int32),
        ("n_gpus", C.c_int32), ("first_device", C.c_int32), ("fused_epilogue", C.c_int32), ("verbose", C.c_int32), ("pd_check_passes", C.c_int32), ("p2p_exchange", C.c_int32),
        ("exchange_timeout_ms", C.c_int32), ("allow_unconverged_init", C.c_int32),
    ]


class Out(C.Structure):
    _fields_ = [
        ("vectors", C.POINTER(C.c_double)), ("standard_vectors", C.POINTER(C.c_double)), ("values", C.POINTER(C.c_double)),
        ("F", C.POINTER(C.c_double)), ("step_a", C.POINTER(C.c_double)), ("backtracks_per_k", C.POINTER(C.c_int32)),
        ("stage_of_k", C.POINTER(C.c_int32)), ("trace_capacity", C.c_int32),
        ("iterations", C.c_int32), ("backtracks", C.c_int32), ("contractions", C.c_int32), ("init_passes", C.c_int32),
        ("init_converged", C.c_int32), ("init_residual", C.c_double), ("seconds_upload", C.c_double),
        ("seconds_init", C.c_double), ("seconds_loop", C.c_double), ("seconds_setup", C.c_double),
        ("seconds_prepare", C.c_double), ("upload_staged", C.c_int32), ("polar_passes_max", C.c_int32),
    ]


class CovCfg(C.Structure):
    """speigencov_cfg of include/speigen_b200.h (constants of R/spEigenCov.R:49,69,73,80-88)."""
    _fields_ = [
        ("max_iter", C.c_int32), ("n_continuation", C.c_int32), ("p1", C.c_double), ("pk", C.c_double),
        ("tol_factor", C.c_double), ("eps_factor", C.c_double), ("shift", C.c_double), ("rank_tol", C.c_double),
        ("psd_tol", C.c_double), ("rho_scale", C.c_double), ("check_symmetric", C.c_int32), ("check_na", C.c_int32),
        ("device", C.c_int32), ("jacobi_max_sweeps", C.c_int32), ("jacobi_tol", C.c_double), ("polish", C.c_int32),
        ("verbose", C.c_int32),
    ]


class CovOut(C.Structure):
    _fields_ = [
        ("vectors", C.POINTER(C.c_double)), ("values", C.POINTER(C.c_double)), ("cov", C.POINTER(C.c_double)),
        ("F", C.POINTER(C.c_double)), ("stage_of_k", C.POINTER(C.c_int32)), ("trace_capacity", C.c_int32),
        ("iterations", C.c_int32), ("eig_sweeps", C.c_int32), ("svd_sweeps", C.c_int32), ("gemms", C.c_int32),
        ("seconds_upload", C.c_double), ("seconds_init", C.c_double), ("seconds_loop", C.c_double),
        ("seconds_gemm", C.c_double),
    ]


class Problem(C.Structure):
    _fields_ = [("m", C.c_int64), ("n", C.c_int64), ("q", C.c_int32), ("is_complex", C.c_int32), ("is_data", C.c_int32),
                ("world", C.c_int32), ("rank", C.c_int32), ("local_gpus", C.c_int32)]


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    """Load libspeigen_b200.so (fails loudly: the product path has no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C sparseeigen_b200/csrc`; there is no CPU fallback")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    L.speigen_last_error.restype = C.c_char_p
    import atexit
    atexit.register(L.speigen_shutdown)
    L.speigen_cfg_default.argtypes = [C.POINTER(Cfg)]
    L.speigen_run_f64.argtypes = [_dp, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_double,
                                  C.POINTER(Cfg), C.POINTER(Out)]
    L.speigen_run_c64.argtypes = L.speigen_run_f64.argtypes
    L.speigen_validate_args.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_double]
    L.speigen_default_d.argtypes = [C.c_int32, _dp]
    L.speigen_schedule.argtypes = [C.POINTER(Cfg), _dp, _dp, _dp]
    L.speigen_stage_constants.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
    L.speigen_shard_range.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.speigen_panel_pitch.argtypes = [C.c_int32]
    L.speigen_panel_pitch.restype = C.c_int32
    L.speigen_matrix_ld.argtypes = [C.c_int64]
    L.speigen_matrix_ld.restype = C.c_int64
    L.speigen_solver_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Problem), C.POINTER(Cfg)]
    L.speigen_solver_destroy.argtypes = [C.c_void_p]
    L.speigen_nccl_unique_id.argtypes = [C.c_void_p]
    L.speigen_solver_init_comm.argtypes = [C.c_void_p, C.c_void_p]
    L.speigen_solver_load_host.argtypes = [C.c_void_p, _dp]
    L.speigen_solver_shard_buffer.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)] + [C.POINTER(C.c_int64)] * 5
    L.speigen_solver_prepare.argtypes = [C.c_void_p]
    L.speigen_solver_init.argtypes = [C.c_void_p, _dp, C.POINTER(Out)]
    L.speigen_solver_run.argtypes = [C.c_void_p, C.c_double, _dp, C.c_double, C.POINTER(Out)]
    L.speigen_solver_contract.argtypes = [C.c_void_p, _dp, C.c_int32, _dp]
    L.speigen_solver_polar.argtypes = [C.c_void_p, _dp, C.c_int32, _dp]
    L.speigen_solver_mm_update.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int32, _dp]
    L.speigen_solver_objective.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    L.speigen_solver_eig_small.argtypes = [C.c_void_p, _dp, C.c_int32, _dp, _dp]
    L.speigen_solver_time_contract.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
    L.speigen_solver_launch_count.argtypes = [C.c_void_p]
    L.speigen_solver_launch_count.restype = C.c_int64
    L.speigen_solver_stream.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.speigen_solver_load_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
    L.speigen_solver_load_device_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
    L.speigen_solver_mm_steps.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, _dp, C.POINTER(C.c_float)]
    L.speigen_solver_sync.argtypes = [C.c_void_p]
    L.speigen_solver_tail_diag.argtypes = [C.c_void_p, _dp]
    L.speigen_solver_selftest_timeout.argtypes = [C.c_void_p, C.c_int32]
    L.speigen_solver_upload_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
    L.speigen_host_half_upload_emulate.argtypes = [_dp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                                   _dp, _dp, _dp]
    L.speigen_host_has_avx2.argtypes = []
    L.speigen_last_call_times.argtypes = [_dp]
    L.speigen_last_upload_bytes.argtypes = []
    L.speigen_last_upload_bytes.restype = C.c_int64
    L.speigen_solver_exchange_mode.argtypes = [C.c_void_p]
    L.speigen_solver_kernel_timing.argtypes = [C.c_void_p, C.c_int32]
    L.speigen_solver_kernel_time.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_int32)]
    _ip = C.POINTER(C.c_int32)
    _fp = C.POINTER(C.c_float)
    L.speigencov_cfg_default.argtypes = [C.POINTER(CovCfg)]
    L.speigencov_run_f64.argtypes = [_dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.POINTER(CovCfg),
                                     C.POINTER(CovOut)]
    L.speigencov_run_c64.argtypes = L.speigencov_run_f64.argtypes
    L.speigencov_validate_args.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_double]
    L.speigencov_schedule.argtypes = [C.POINTER(CovCfg), _dp, _dp, _dp]
    L.speigencov_eigval_algo.argtypes = [_dp, C.c_int64, C.c_double, C.c_int32, C.c_int32, _dp]
    L.speigen_dense_gemm.argtypes = [C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int, _dp, C.c_int64, C.c_int, _dp,
                                     C.c_int64, C.c_double, C.c_int, C.c_int, _fp]
    L.speigen_dense_gemm_time.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32, _fp]
    L.speigen_dense_polar.argtypes = [_dp, C.c_int64, _dp, C.c_int32, C.c_int, _ip]
    L.speigen_dense_eigh.argtypes = [_dp, C.c_int64, _dp, _dp, C.c_int, _ip]
    L.speigen_solver_load_cov_of_data.argtypes = [C.c_void_p, _dp, C.c_int64]
    L.speigen_solver_load_cov_of_data_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    L.speigen_run_cov_of_data_f64.argtypes = [_dp, C.c_int64, C.c_int64, C.c_double, C.c_double, _dp, _dp, C.c_double,
                                              C.POINTER(Cfg), C.POINTER(Out)]
    L.speigen_cov_f64.argtypes = [_dp, C.c_int64, C.c_int64, _dp, C.c_int, _fp]
    sz = (C.c_int32 * 3)()
    L.speigen_struct_sizes(C.byref(sz, 0), C.byref(sz, 4), C.byref(sz, 8))
    if (sz[0], sz[1], sz[2]) != (C.sizeof(Cfg), C.sizeof(Out), C.sizeof(Problem)):
        raise ImportError(f"{path}: struct layout mismatch with include/speigen_b200.h "
                          f"(library {tuple(sz)} vs binding {(C.sizeof(Cfg), C.sizeof(Out), C.sizeof(Problem))}); rebuild the library")
    sz2 = (C.c_int32 * 2)()
    L.speigencov_struct_sizes(C.byref(sz2, 0), C.byref(sz2, 4))
    if (sz2[0], sz2[1]) != (C.sizeof(CovCfg), C.sizeof(CovOut)):
        raise ImportError(f"{path}: speigencov struct layout mismatch (library {tuple(sz2)} vs binding "
                          f"{(C.sizeof(CovCfg), C.sizeof(CovOut))}); rebuild the library")
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise SpEigenError(rc, lib().speigen_last_error().decode("utf-8", "replace"))


def default_cfg(**overrides) -> Cfg:
    cfg = Cfg()
    lib().speigen_cfg_default(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown cfg field {k}")
        setattr(cfg, k, v)
    return cfg


def schedule(cfg: Optional[Cfg] = None):
    """(p, eps, tol) of R/spEigen.R:94-102."""
    cfg = cfg or default_cfg()
    n = cfg.n_continuation + 1
    p, e, t = (np.empty(n) for _ in range(3))
    _check(lib().speigen_schedule(C.byref(cfg), _ptr(p), _ptr(e), _ptr(t)))
    return p, e, t


def default_d(q: int) -> np.ndarray:
    d = np.empty(q)
    _check(lib().speigen_default_d(q, _ptr(d)))
    return d


def shard_range(n: int, nparts: int, part: int):
    lo, hi = C.c_int64(), C.c_int64()
    _check(lib().speigen_shard_range(n, nparts, part, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def panel_pitch(ncols_real: int) -> int:
    return lib().speigen_panel_pitch(ncols_real)


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(_dp)


def _as_colmajor(X: np.ndarray, cplx: bool) -> np.ndarray:
    """Column-major float64 / complex128 buffer exactly as R stores a matrix."""
    return np.asfortranarray(X, dtype=np.complex128 if cplx else np.float64)


def _is_na(v) -> bool:
    if v is None:
        return True
    try:
        return bool(np.isnan(v))
    except TypeError:
        return False


def _trace_capacity(cfg) -> int:
    """k can pass max_iter: reaching it ends a continuation stage, not the schedule (R/spEigen.R:151), and every remaining stage
    still runs two iterations."""
    return max(cfg.max_iter, 1) + 2 * (cfg.n_continuation + 1) + 8


class _OutBuffers:
    def __init__(self, m, q, cplx, trace_capacity=1000):
        dt = np.complex128 if cplx else np.float64
        self.vectors = np.zeros((m, q), dtype=dt, order="F")
        self.standard_vectors = np.zeros((m, q), dtype=dt, order="F")
        self.values = np.zeros(q)
        self.F = np.zeros(trace_capacity)
        self.step_a = np.zeros(trace_capacity)
        self.bt = np.zeros(trace_capacity, dtype=np.int32)
        self.stage = np.zeros(trace_capacity, dtype=np.int32)
        o = Out()
        o.vectors = self.vectors.ctypes.data_as(_dp)
        o.standard_vectors = self.standard_vectors.ctypes.data_as(_dp)
        o.values = _ptr(self.values)
        o.F = _ptr(self.F)
        o.step_a = _ptr(self.step_a)
        o.backtracks_per_k = self.bt.ctypes.data_as(C.POINTER(C.c_int32))
        o.stage_of_k = self.stage.ctypes.data_as(C.POINTER(C.c_int32))
        o.trace_capacity = trace_capacity
        self.c = o

    def result(self):
        k = min(self.c.iterations, len(self.F))
        info = dict(iterations=self.c.iterations, backtracks=self.c.backtracks, contractions=self.c.contractions,
                    init_passes=self.c.init_passes, init_converged=bool(self.c.init_converged),
                    init_residual=self.c.init_residual, seconds_upload=self.c.seconds_upload,
                    seconds_init=self.c.seconds_init, seconds_loop=self.c.seconds_loop, seconds_setup=self.c.seconds_setup,
                    seconds_prepare=self.c.seconds_prepare, upload_staged=bool(self.c.upload_staged & 1), upload_half=bool(self.c.upload_staged & 2),
                    polar_passes_max=self.c.polar_passes_max,
                    F=self.F[:k].copy(), step_a=self.step_a[:k].copy(), backtracks_per_k=self.bt[:k].copy(),
                    stage_of_k=self.stage[:k].copy())
        return dict(vectors=self.vectors, standard_vectors=self.standard_vectors, values=self.values, info=info)


def spEigen(X, q=1, rho=0.5, data=False, d=None, V=None, thres=1e-9, *, n_gpus=1, cfg: Optional[Cfg] = None,
            cov_of_data=False, **cfg_overrides):
    """Sparse (orthogonal) eigenvectors of a covariance or data matrix — drop-in for R/spEigen.R:46.

    Returns dict(vectors, standard_vectors, values) as the reference's list (R/spEigen.R:162) plus an
    'info' side channel (objective trajectory F_v, which the reference computes but does not return).
    `d=None` / `V=None` play the role of R's NA defaults.

    cov_of_data=True evaluates the callers' `spEigen(cov(X), q, ...)` (README.Rmd:108) with X the n x m data matrix:
    cov(X) is formed by the FP64 tensor-core GEMM straight into the device-resident shards, so the
    m x m matrix never exists on the host.  Real X only."""
    X = np.asarray(X)
    if cov_of_data:
        if data:
            raise SpEigenError(23, "cov_of_data=True means spEigen(cov(X), ...): use data=False")
        if np.iscomplexobj(X) or X.ndim != 2:
            raise SpEigenError(25, "cov_of_data needs a real n x m matrix")
        n, m = X.shape
        if m == 1:
            raise SpEigenError(1, "Data is univariate!")
        if _is_na(q) or _is_na(rho):
            raise SpEigenError(2, "This function cannot handle NAs.")
        L = lib()
        _check(L.speigen_validate_args(m, m, float(q), float(rho)))
        q = int(q)
        Xf = _as_colmajor(X, False)
        cfg = cfg or default_cfg()
        cfg.n_gpus = n_gpus
        for k, v in cfg_overrides.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown cfg field {k}")
            setattr(cfg, k, v)
        dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
        VV = None if V is None else _as_colmajor(np.asarray(V), False)
        ob = _OutBuffers(m, q, False, trace_capacity=_trace_capacity(cfg))
        _check(L.speigen_run_cov_of_data_f64(Xf.ctypes.data_as(_dp), n, m, float(q), float(rho), _ptr(dd),
                                             None if VV is None else VV.ctypes.data_as(_dp), float(thres),
                                             C.byref(cfg), C.byref(ob.c)))
        return ob.result()
    if X.ndim == 1:                      # as.matrix(<vector>) is n x 1   (R/spEigen.R:50)
        X = X.reshape(-1, 1)
    if X.ndim != 2:
        raise SpEigenError(23, "X must be a matrix")
    nrow, m = X.shape
    if m == 1:
        raise SpEigenError(1, "Data is univariate!")                                   # :52
    if _is_na(q) or _is_na(rho):
        raise SpEigenError(2, "This function cannot handle NAs.")                      # :53
    # anyNA(X) (:53) is evaluated by the device scan in speigen_solver_prepare (one streamed pass, no host copy)
    L = lib()
    _check(L.speigen_validate_args(nrow, m, float(q), float(rho)))                     # :54-55
    q = int(q)
    cplx = np.iscomplexobj(X)
    Xf = _as_colmajor(X, cplx)
    cfg = cfg or default_cfg()
    cfg.n_gpus = n_gpus
    for k, v in cfg_overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown cfg field {k}")
        setattr(cfg, k, v)
    dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
    if dd is not None and dd.size != q:
        raise SpEigenError(23, "d must have length q")
    VV = None
    if V is not None:
        VV = _as_colmajor(np.asarray(V), cplx)
        if VV.shape != (m, q):
            raise SpEigenError(23, "V must be m x q")
    ob = _OutBuffers(m, q, cplx, trace_capacity=_trace_capacity(cfg))
    fn = L.speigen_run_c64 if cplx else L.speigen_run_f64
    rc = fn(Xf.ctypes.data_as(_dp), nrow, m, 1 if data else 0, float(q), float(rho), _ptr(dd),
            None if VV is None else VV.ctypes.data_as(_dp), float(thres), C.byref(cfg), C.byref(ob.c))
    _check(rc)
    return ob.result()


def default_cov_cfg(**overrides) -> CovCfg:
    cfg = CovCfg()
    lib().speigencov_cfg_default(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown cfg field {k}")
        setattr(cfg, k, v)
    return cfg


def cov_schedule(cfg: Optional[CovCfg] = None):
    """(p, eps, tol) of R/spEigenCov.R:80-88."""
    cfg = cfg or default_cov_cfg()
    n = cfg.n_continuation + 1
    p, e, t = (np.empty(n) for _ in range(3))
    _check(lib().speigencov_schedule(C.byref(cfg), _ptr(p), _ptr(e), _ptr(t)))
    return p, e, t


def eigval_algo(g, x, q, strict_r=False) -> np.ndarray:
    """eigvalAlgo(g, x, q) of R/spEigenCov.R:164-338 (host-side part of the spEigenCov path)."""
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.empty_like(g)
    _check(lib().speigencov_eigval_algo(_ptr(g), g.size, float(x), int(q), 1 if strict_r else 0, _ptr(out)))
    return out


def spEigenCov(S, q=1, rho=0.5, thres=1e-9, *, cfg: Optional[CovCfg] = None, want_cov=True, **cfg_overrides):
    """Covariance estimation with sparse eigenvectors — drop-in for R/spEigenCov.R:48.

    Returns dict(vectors m x m, values m, cov m x m) as the reference's list (R/spEigenCov.R:135) plus an 'info' side
    channel (objective trajectory, sweep counts, timings).  Real or complex Hermitian S."""
    S = np.asarray(S)
    if S.ndim == 0:
        S = S.reshape(1, 1)
    if S.ndim == 1:                      # as.matrix(<vector>) is n x 1   (R/spEigenCov.R:52)
        S = S.reshape(-1, 1)
    if S.ndim != 2:
        raise SpEigenError(23, "S must be a matrix")
    nrow, m = S.shape
    L = lib()
    qq = float("nan") if _is_na(q) else float(q)
    rr = float("nan") if _is_na(rho) else float(rho)
    _check(L.speigencov_validate_args(nrow, m, qq, rr))                               # :54-60
    cplx = np.iscomplexobj(S)
    Sf = _as_colmajor(S, cplx)
    cfg = cfg or default_cov_cfg()
    for k, v in cfg_overrides.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown cfg field {k}")
        setattr(cfg, k, v)
    cap = max(cfg.max_iter, 1)
    dt = np.complex128 if cplx else np.float64
    vectors = np.zeros((m, m), dtype=dt, order="F")
    values = np.zeros(m)
    covm = np.zeros((m, m), dtype=dt, order="F") if want_cov else None
    F = np.zeros(cap)
    stage = np.zeros(cap, dtype=np.int32)
    o = CovOut()
    o.vectors = vectors.ctypes.data_as(_dp)
    o.values = _ptr(values)
    o.cov = None if covm is None else covm.ctypes.data_as(_dp)
    o.F = _ptr(F)
    o.stage_of_k = stage.ctypes.data_as(C.POINTER(C.c_int32))
    o.trace_capacity = cap
    fn = L.speigencov_run_c64 if cplx else L.speigencov_run_f64
    _check(fn(Sf.ctypes.data_as(_dp), nrow, m, qq, rr, float(thres), C.byref(cfg), C.byref(o)))
    k = min(o.iterations, cap)
    info = dict(iterations=o.iterations, eig_sweeps=o.eig_sweeps, svd_sweeps=o.svd_sweeps, gemms=o.gemms,
                seconds_upload=o.seconds_upload, seconds_init=o.seconds_init, seconds_loop=o.seconds_loop,
                seconds_gemm=o.seconds_gemm, F=F[:k].copy(), stage_of_k=stage[:k].copy())
    return dict(vectors=vectors, values=values, cov=covm, info=info)


def dense_gemm(A, B, a_kmajor, b_kmajor, alpha=1.0, symmetric=False, device=0, return_ms=False):
    """C = alpha * op(A) op(B) through the DMMA GEMM kernel.  a_kmajor: A is K x M (op(A) = A^T) else M x K;
    b_kmajor: B is K x N else N x K (op(B) = B^T)."""
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    K, M = (A.shape if a_kmajor else A.shape[::-1])
    Kb, N = (B.shape if b_kmajor else B.shape[::-1])
    if K != Kb:
        raise SpEigenError(23, "inner dimensions differ")
    Cm = np.zeros((M, N), order="F")
    ms = C.c_float(0.0)
    _check(lib().speigen_dense_gemm(M, N, K, _ptr(A), A.shape[0], 1 if a_kmajor else 0, _ptr(B), B.shape[0],
                                    1 if b_kmajor else 0, _ptr(Cm), M, float(alpha), 1 if symmetric else 0, device,
                                    C.byref(ms)))
    return (Cm, ms.value) if return_ms else Cm


def dense_gemm_time(M, N, K, a_kmajor=True, b_kmajor=True, symmetric=False, reps=5, device=0) -> np.ndarray:
    ms = (C.c_float * reps)()
    _check(lib().speigen_dense_gemm_time(M, N, K, int(a_kmajor), int(b_kmajor), int(symmetric), device, reps, ms))
    return np.array(ms[:], dtype=np.float64)


def dense_polar(B, polish=True, device=0):
    """(U, sweeps): U = u v^T of svd(B) (R/spEigenCov.R:157-159) by the block one-sided Jacobi kernels."""
    B = np.asfortranarray(B, dtype=np.float64)
    m = B.shape[0]
    U = np.zeros((m, m), order="F")
    sw = C.c_int32(0)
    _check(lib().speigen_dense_polar(_ptr(B), m, _ptr(U), 1 if polish else 0, device, C.byref(sw)))
    return U, sw.value


def dense_eigh(S, device=0):
    """(values decreasing, vectors, sweeps) as eigen(S) (R/spEigenCov.R:62)."""
    S = np.asfortranarray(S, dtype=np.float64)
    m = S.shape[0]
    V = np.zeros((m, m), order="F")
    w = np.zeros(m)
    sw = C.c_int32(0)
    _check(lib().speigen_dense_eigh(_ptr(S), m, _ptr(V), _ptr(w), device, C.byref(sw)))
    return w, V, sw.value


def cov(X, device=0, return_ms=False):
    """cov(X) (column-centred, / (n-1)) formed on the device by the symmetric DMMA GEMM."""
    X = np.asfortranarray(X, dtype=np.float64)
    n, m = X.shape
    S = np.zeros((m, m), order="F")
    ms = C.c_float(0.0)
    _check(lib().speigen_cov_f64(_ptr(X), n, m, _ptr(S), device, C.byref(ms)))
    return (S, ms.value) if return_ms else S


class Solver:
    """Resident solver handle (speigen_solver_* of the C ABI): kernel-level entry points for the parity
    tests and the device-resident benchmark path."""

    def __init__(self, m, q, *, n=0, is_complex=False, is_data=False, world=1, rank=0, local_gpus=1,
                 cfg: Optional[Cfg] = None, **cfg_overrides):
        self.L = lib()
        self.cfg = cfg or default_cfg(**cfg_overrides)
        self.m, self.n, self.q = int(m), int(n), int(q)
        self.cplx, self.data = bool(is_complex), bool(is_data)
        self.prob = Problem(m=self.m, n=self.n, q=self.q, is_complex=int(self.cplx), is_data=int(self.data), world=world,
                            rank=rank, local_gpus=local_gpus)
        self.h = C.c_void_p()
        _check(self.L.speigen_solver_create(C.byref(self.h), C.byref(self.prob), C.byref(self.cfg)))
        self._keep = []

    def close(self):
        if self.h:
            self.L.speigen_solver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def dtype(self):
        return np.complex128 if self.cplx else np.float64

    def init_comm(self, unique_id: Optional[bytes] = None):
        if unique_id is None:
            _check(self.L.speigen_solver_init_comm(self.h, None))
        else:
            buf = C.create_string_buffer(unique_id, 128)
            _check(self.L.speigen_solver_init_comm(self.h, buf))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().speigen_nccl_unique_id(buf))
        return buf.raw

    def load(self, X):
        Xf = _as_colmajor(np.asarray(X), self.cplx)
        _check(self.L.speigen_solver_load_host(self.h, Xf.ctypes.data_as(_dp)))

    def load_cov_of_data(self, X):
        """S = cov(X) formed on the device(s) straight into the shard buffers (covariance-mode solver)."""
        Xf = _as_colmajor(np.asarray(X), False)
        _check(self.L.speigen_solver_load_cov_of_data(self.h, Xf.ctypes.data_as(_dp), Xf.shape[0]))

    def load_cov_of_data_device(self, dev_ptr: int, n: int, ldx: int):
        _check(self.L.speigen_solver_load_cov_of_data_device(self.h, C.c_void_p(dev_ptr), n, ldx))

    def shard_buffer(self, local_idx=0):
        p = C.c_void_p()
        v = [C.c_int64() for _ in range(5)]
        _check(self.L.speigen_solver_shard_buffer(self.h, local_idx, C.byref(p), *[C.byref(x) for x in v]))
        return dict(ptr=p.value, rows=v[0].value, row_len=v[1].value, ld=v[2].value, lo=v[3].value, hi=v[4].value)

    def prepare(self):
        _check(self.L.speigen_solver_prepare(self.h))

    def init(self, V0=None):
        ob = _OutBuffers(self.m, self.q, self.cplx, 1)
        VV = None if V0 is None else _as_colmajor(np.asarray(V0), self.cplx)
        _check(self.L.speigen_solver_init(self.h, None if VV is None else VV.ctypes.data_as(_dp), C.byref(ob.c)))
        return dict(init_passes=ob.c.init_passes, init_converged=bool(ob.c.init_converged),
                    init_residual=ob.c.init_residual, seconds_init=ob.c.seconds_init)

    def run(self, rho=0.5, d=None, thres=1e-9):
        ob = _OutBuffers(self.m, self.q, self.cplx, _trace_capacity(self.cfg))
        dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
        _check(self.L.speigen_solver_run(self.h, float(rho), _ptr(dd), float(thres), C.byref(ob.c)))
        return ob.result()

    def contract(self, U):
        U = _as_colmajor(np.asarray(U), self.cplx)
        Y = np.zeros_like(U, order="F")
        _check(self.L.speigen_solver_contract(self.h, U.ctypes.data_as(_dp), U.shape[1], Y.ctypes.data_as(_dp)))
        return Y

    def polar(self, A):
        A = _as_colmajor(np.asarray(A), self.cplx)
        U = np.zeros_like(A, order="F")
        _check(self.L.speigen_solver_polar(self.h, A.ctypes.data_as(_dp), A.shape[1], U.ctypes.data_as(_dp)))
        return U

    def mm_update(self, V, d, rho_vec, p, eps, fused=True):
        V = _as_colmajor(np.asarray(V), self.cplx)
        V1 = np.zeros_like(V, order="F")
        d = np.ascontiguousarray(d, dtype=np.float64)
        r = np.ascontiguousarray(rho_vec, dtype=np.float64)
        _check(self.L.speigen_solver_mm_update(self.h, V.ctypes.data_as(_dp), _ptr(d), _ptr(r), float(p), float(eps),
                                               1 if fused else 0, V1.ctypes.data_as(_dp)))
        return V1

    def objective(self, V, d, rho_vec, p, eps) -> float:
        V = _as_colmajor(np.asarray(V), self.cplx)
        d = np.ascontiguousarray(d, dtype=np.float64)
        r = np.ascontiguousarray(rho_vec, dtype=np.float64)
        F = C.c_double()
        _check(self.L.speigen_solver_objective(self.h, V.ctypes.data_as(_dp), _ptr(d), _ptr(r), float(p), float(eps),
                                               C.byref(F)))
        return F.value

    def eig_small(self, Cm):
        Cm = _as_colmajor(np.asarray(Cm), self.cplx)
        n = Cm.shape[0]
        W = np.zeros_like(Cm, order="F")
        ev = np.zeros(n)
        _check(self.L.speigen_solver_eig_small(self.h, Cm.ctypes.data_as(_dp), n, W.ctypes.data_as(_dp), _ptr(ev)))
        return ev, W

    def time_contract(self, ncols, reps, flush_l2=True):
        ms = (C.c_float * reps)()
        _check(self.L.speigen_solver_time_contract(self.h, ncols, reps, 1 if flush_l2 else 0, ms))
        return np.array(ms[:], dtype=np.float64)

    def load_device(self, dev_ptr: int, src_ld: int, local_idx=0):
        _check(self.L.speigen_solver_load_device(self.h, local_idx, C.c_void_p(dev_ptr), src_ld))

    def load_device_rows(self, row0: int, nrows: int, dev_ptr: int, src_ld: int, local_idx=0):
        _check(self.L.speigen_solver_load_device_rows(self.h, local_idx, row0, nrows, C.c_void_p(dev_ptr), src_ld))

    def mm_steps(self, nsteps, stage=0, rho=0.5, d=None, timed=False):
        """nsteps MM updates on the resident iterate; timed=True returns the CUDA-event elapsed ms."""
        dd = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
        ms = C.c_float(0.0)
        _check(self.L.speigen_solver_mm_steps(self.h, nsteps, stage, float(rho), _ptr(dd), C.byref(ms) if timed else None))
        return ms.value if timed else None

    def exchange_mode(self) -> str:
        return {0: "none", 1: "nccl", 2: "p2p-fused"}[self.L.speigen_solver_exchange_mode(self.h)]

    def sync(self):
        _check(self.L.speigen_solver_sync(self.h))

    def tail_diag(self):
        """Diagnostics of the last fused tail kernel on device 0 (passes, Jacobi sweeps, CTA 0's phase clock in microseconds)."""
        out = np.zeros(48)
        _check(self.L.speigen_solver_tail_diag(self.h, _ptr(out)))
        n = int(out[7])
        return dict(a=out[0], nR=out[1], nU=out[2], passes=int(out[3]), sweeps=int(out[4]),
                    clock_us=[float(x) for x in np.round(out[8:8 + n] * 1e-3, 2)],
                    jacobi_us=dict(setup=float((out[42] - out[41]) * 1e-3), sweeps=float((out[40] - out[42]) * 1e-3)))

    def kernel_timing(self, enable=True):
        _check(self.L.speigen_solver_kernel_timing(self.h, 1 if enable else 0))

    def kernel_time(self):
        ms, n = C.c_double(), C.c_int32()
        _check(self.L.speigen_solver_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.L.speigen_solver_launch_count(self.h))
