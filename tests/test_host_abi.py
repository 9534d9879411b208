"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
and its host-only logic (validation, schedule, defaults, sharding) agrees with the oracle.  No compute
calls are made without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import sparseeigen_b200 as sb
from oracle import speigen_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "speigen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(speigen(?:cov)?_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = sb.lib()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/speigen_b200.h but not exported"
    assert L.speigen_abi_version() == 1


def test_r_glue_uses_only_exported_symbols():
    """integration/r-package/src/speigen_glue.c (the .Call stub of INTEGRATION.md) binds only exported ABI symbols."""
    src = open(os.path.join(ROOT, "integration", "r-package", "src", "speigen_glue.c")).read()
    used = set(re.findall(r"\b(speigen(?:cov)?_(?:run_f64|run_c64|run_cov_of_data_f64|cfg_default|last_error|shutdown))\s*\(", src))
    assert {"speigen_run_f64", "speigen_run_c64", "speigen_cfg_default", "speigen_last_error", "speigen_shutdown",
            "speigencov_run_f64", "speigencov_cfg_default", "speigen_run_cov_of_data_f64"} <= used
    L = sb.lib()
    for s in used:
        assert hasattr(L, s)


def test_r_glue_type_checks_against_the_header():
    """`gcc -fsyntax-only` of the .Call glue against include/speigen_b200.h and declaration-only stand-ins for the R
    headers (tests/stubs/r_headers; R itself is not installed in the image): argument counts and types of every ABI call
    the glue makes are checked by the compiler."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "stubs", "r_headers"),
                        "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "r-package", "src", "speigen_glue.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_header_compiles_as_plain_c():
    """include/speigen_b200.h is a C header (no C++/torch types in any signature)."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "speigen_b200.h"\nint main(void) { speigen_cfg c; speigencov_cfg d; (void)c; (void)d; return 0; }\n')
        path = f.name
    try:
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I",
                            os.path.join(ROOT, "include"), path], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    finally:
        os.unlink(path)


def test_struct_layouts_match_the_header():
    L = sb.lib()
    sz = (ctypes.c_int32 * 3)()
    assert L.speigen_struct_sizes(ctypes.byref(sz, 0), ctypes.byref(sz, 4), ctypes.byref(sz, 8)) == 0
    assert tuple(sz) == (ctypes.sizeof(sb.Cfg), ctypes.sizeof(sb.Out), ctypes.sizeof(sb.Problem))
    cfg = sb.default_cfg()
    assert (cfg.max_iter, cfg.n_continuation, cfg.max_backtracks) == (1000, 10, 200)     # R/spEigen.R:47,94
    assert (cfg.p1, cfg.pk, cfg.tol_factor, cfg.backtrack_slack) == (1.0, 7.0, 1e-2, 1e-9)   # :95-96,101,140
    assert (cfg.rank_tol, cfg.pd_tol) == (1e-9, -1e-10) and cfg.p2p_exchange == 1 and cfg.pd_check_passes == 4


def test_validate_args_matches_reference_conditions():
    L = sb.lib()
    assert L.speigen_validate_args(15, 10, 3.0, 0.5) == 0
    assert L.speigen_validate_args(15, 1, 3.0, 0.5) == 1          # univariate      R/spEigen.R:52
    assert L.speigen_validate_args(15, 10, float("nan"), 0.5) == 2   # NA           :53
    assert L.speigen_validate_args(15, 10, 3.0, float("nan")) == 2
    for q in (-1.0, 0.0, 0.5):
        assert L.speigen_validate_args(15, 10, q, 0.5) == 3       # q               :54
    assert L.speigen_validate_args(15, 10, 3.0, 0.0) == 4         # rho <= 0        :55
    assert b"positive" in L.speigen_last_error()


def test_schedule_defaults_stage_constants_match_oracle():
    p, e, t = sb.schedule()
    po, eo, to = o.schedule()
    assert np.allclose(p, po, rtol=1e-15) and np.allclose(e, eo, rtol=1e-15) and np.allclose(t, to, rtol=1e-15)
    for q in (1, 2, 3, 20):
        assert np.allclose(sb.default_d(q), o.default_d(q), rtol=1e-15)
    L = sb.lib()
    c1, c2, w0 = (ctypes.c_double() for _ in range(3))
    for pp, ee in zip(po, eo):
        L.speigen_stage_constants(pp, ee, ctypes.byref(c1), ctypes.byref(c2), ctypes.byref(w0))
        r1, r2, r0 = o.stage_constants(pp, ee)
        assert np.isclose(c1.value, r1, rtol=1e-15) and np.isclose(c2.value, r2, rtol=1e-15) and np.isclose(w0.value, r0, rtol=1e-15)


@pytest.mark.parametrize("n,parts", [(50000, 8), (50000, 1), (10, 4), (7, 8), (200000, 3)])
def test_shard_ranges_partition(n, parts):
    prev = 0
    sizes = []
    for r in range(parts):
        lo, hi = sb.shard_range(n, parts, r)
        assert lo == prev and hi >= lo
        prev = hi
        sizes.append(hi - lo)
    assert prev == n
    assert max(sizes) == -(-n // parts)            # equal blocks of ceil(n/parts): NCCL allgather needs equal counts


def test_panel_pitch_and_ld():
    for ncr in range(1, 65):
        p = sb.panel_pitch(ncr)
        assert p >= ncr and p % 8 == 4             # bank-conflict-free DMMA B-fragment reads, 32-byte aligned rows
    L = sb.lib()
    assert L.speigen_matrix_ld(50000) == 50000 and L.speigen_matrix_ld(50001) == 50016


def test_python_wrapper_error_behaviour_without_gpu():
    X = np.random.default_rng(0).standard_normal((15, 10))
    with pytest.raises(sb.SpEigenError): sb.spEigen(X[:, 0], 3, 0.5, data=True)
    Xna = X.copy(); Xna[0, 0] = np.nan
    with pytest.raises(sb.SpEigenError): sb.spEigen(Xna, 3, 0.5, data=True)
    with pytest.raises(sb.SpEigenError): sb.spEigen(X, None, 0.5, data=True)
    with pytest.raises(sb.SpEigenError): sb.spEigen(X, 3, float("nan"), data=True)
    for q in (-1, 0, 0.5):
        with pytest.raises(sb.SpEigenError): sb.spEigen(X, q, 0.5, data=True)
    with pytest.raises(sb.SpEigenError): sb.spEigen(X, 3, 0.0, data=True)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through the oracle)."""
    if sb.lib().speigen_device_count() > 0:
        pytest.skip("a GPU is present")
    X = np.random.default_rng(0).standard_normal((15, 10))
    with pytest.raises(sb.SpEigenError) as ei:
        sb.spEigen(X, 3, 0.5, data=True)
    assert ei.value.code == 22 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sparseeigen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_bench_uses_the_oracle_only_in_its_cpu_legs():
    """bench.py may execute oracle/ only as the CPU baseline / reference arm (never in the measured GPU path)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"cpu_mm_update_time", "bench_speigencov"}      # the latter imports it inside its `if not args.no_cpu` leg only
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and node.module == "oracle":
                assert fn.name in allowed, f"bench.py:{fn.name} imports oracle"
    body = src[src.index("def bench_speigencov"):src.index("def main")]
    assert body.index("from oracle import") > body.index("if not args.no_cpu"), "oracle import outside the cpu_baseline leg"
    for n in tree.body:                                          # no module-level import of the oracle
        assert not (isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("oracle"))


def test_synthetic_generator_matches_the_oracles():
    """The package's benchmark generator draws the same bits as the generator the parity tests use."""
    from sparseeigen_b200 import synth
    for cplx in (False, True):
        a = synth.spiked_samples(50, 20, 3, 0.1, seed=7, cplx=cplx)
        b = o.spiked_samples(50, 20, 3, 0.1, seed=7, cplx=cplx)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_tools_and_entry_points_parse():
    """Every script under tools/, bench.py and __graft_entry__.py is syntactically valid (they only run on the GPU box)."""
    import ast
    import glob
    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) >= 10
    for f in files:
        ast.parse(open(f).read(), filename=f)


# ---- Hermitian-half upload: host side (host_pack.h) — the complete procedure emulated in host memory, no GPU -------------
def _emulate_half(S, world, block, slot, simd):
    L = sb.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    Sf = np.asfortranarray(S)
    A = np.full_like(Sf, np.nan)
    diff, sabs = ctypes.c_double(), ctypes.c_double()
    rc = L.speigen_host_half_upload_emulate(Sf.ctypes.data_as(dp), Sf.shape[0], int(np.iscomplexobj(Sf)), world, block, slot,
                                            simd, A.ctypes.data_as(dp), ctypes.byref(diff), ctypes.byref(sabs))
    assert rc == 0
    return A, diff.value, sabs.value


@pytest.mark.parametrize("cplx", [False, True])
def test_half_upload_plan_covers_every_pair_exactly_once(cplx):
    """Every element is either staged ("uploaded") or rebuilt from its transposed partner, for any sharding, super-tile edge
    and staging-slot size; the isSymmetric sums (R/spEigen.R:75) visit every pair exactly once."""
    rng = np.random.default_rng(0)
    for m in (5, 37, 130, 301):
        Z = rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0)
        S = Z + Z.conj().T
        tot = np.abs(S).sum()
        for world in (1, 2, 3, 8):
            for block in (1, 7, 32, 100, 2048):
                for slot in (64, 4096, 8 << 20):
                    for simd in (0, 1, 2):        # 2: pack without compare, sums in a second pass (the one-shot call's deferred flow)
                        A, d, s = _emulate_half(S, world, block, slot, simd)
                        assert np.array_equal(A, S), (m, world, block, slot)
                        assert d == 0.0 and abs(s - tot) < 1e-11 * tot
        for simd in (0, 1, 2):                    # a non-Hermitian matrix: the sums are numpy's
            A, d, s = _emulate_half(Z, 2, 32, 4096, simd)
            dref, sref = np.abs(Z - Z.conj().T).sum(), np.abs(Z).sum()
            assert abs(d - dref) < 1e-11 * dref and abs(s - sref) < 1e-11 * sref
        for (i, j) in ((0, m - 1), (m - 1, 0), (m // 2, m // 2)):     # anyNA (:53): also in a tile that stays on the host
            T = S.copy()
            T[i, j] = np.nan
            assert np.isnan(_emulate_half(T, 2, 32, 4096, 1)[2])


def test_half_upload_moves_about_half_of_the_matrix():
    """With the default 2048 super-tiles, m = 50 000 uploads 13 diagonal tiles + one of each pair: ~52 % of the bytes, and
    every GPU of a row-sharded run carries about the same share."""
    blk, m = 2048, 50000
    nb = -(-m // blk)
    edge = [min(blk, m - i * blk) for i in range(nb)]
    up = np.zeros((nb, nb))
    for i in range(nb):
        for j in range(nb):
            if i == j or ((i + j) % 2 == 0 and i < j) or ((i + j) % 2 == 1 and i > j):
                up[i, j] = edge[i] * edge[j]
    assert 0.5 < up.sum() / m ** 2 < 0.53
    rows = up.sum(axis=1) / (np.array(edge) * m)
    assert rows.min() > 0.45 and rows.max() < 0.60


# ---- the bench contract: every committed bench line of the round carries the keys the driver and the judge read ---------------
@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n2.json", "r02_bench_n4.json", "r02_bench_n8.json",
                                  "r02_bench_n1_final.json", "r02_bench_n2_final.json"])
def test_committed_bench_lines_follow_the_contract(name):
    import json
    path = os.path.join(ROOT, "profiles", name)
    line = json.load(open(path))
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in line, (name, k)
    assert line["metric"].split(" (")[0] in base["metric"] and "m=50k cov, q=20" in line["metric"]
    assert line["dtype"] == "f64" and line["higher_is_better"] is True and line["vs_baseline"] is None and line["warmup"] >= 3
    assert "workload" in line["config"] and "l2" in line["config"]
    assert abs(line["value"] - 1e3 / line["ms_per_step"]) < 1e-6 * line["value"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    # achieved = algorithmic bytes per launch (8 m^2 / P + 16 m q) / the measured launch time
    m, q, P = 50000, 20, line["n_gpus"]
    alg = 8.0 * m * (-(-m // P)) + 16.0 * m * q
    assert abs(r["algorithmic_bytes_per_launch"] - alg) < 1e-3 * alg
    assert abs(r["achieved"] - alg / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-2 * r["achieved"]
    assert r["ms_per_launch"] <= line["ms_per_step"]
    e = line["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, (name, k)
    assert e["host_memory"] == "pageable" and e["n_gpus"] == P
    assert e["value"] < line["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if P == 1:
        c = line["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == line["unit"] and "sample" in c
    else:
        assert line["parity_check"]["ok"] is True and line["parity_check"]["replicas_identical"] is True
