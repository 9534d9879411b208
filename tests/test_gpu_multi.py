"""Multi-GPU parity (-m gpu, needs >= 2 visible B200s; skipped otherwise): the row-sharded covariance path
(NCCL allgather of row blocks) and the sample-sharded data path (allgather of partial panels + rank-ordered
sum) must reproduce the single-GPU results."""
import numpy as np
import pytest

import sparseeigen_b200 as sb
from oracle import speigen_oracle as o
from conftest import colwise_inner

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return sb.lib().speigen_device_count()
    except Exception:
        return 0


need2 = pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")


@need2
@pytest.mark.parametrize("P", [2, 4, 8])
def test_contract_sharded_matches_single(P):
    if _ngpu() < P:
        pytest.skip(f"needs {P} GPUs")
    rng = np.random.default_rng(5)
    m, q = 1037, 20
    A = rng.standard_normal((m, m)); S = (A + A.T) / 2
    U = rng.standard_normal((m, q))
    Ys = []
    for p2p in (1, 0):       # fused peer-store exchange vs NCCL allgather: same bits
        with sb.Solver(m, q, world=P, local_gpus=P, p2p_exchange=p2p) as s:
            s.init_comm(None)
            s.load(S)
            Ys.append(s.contract(U))
            Ys.append(s.contract(U))
    assert np.linalg.norm(Ys[0] - S @ U) / np.linalg.norm(S @ U) < 1e-13
    assert np.array_equal(Ys[0], Ys[1]) and np.array_equal(Ys[0], Ys[2]) and np.array_equal(Ys[0], Ys[3])
    n = 911
    X = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    Uc = rng.standard_normal((m, 5)) + 1j * rng.standard_normal((m, 5))
    with sb.Solver(m, 5, n=n, is_complex=True, is_data=True, world=P, local_gpus=P) as s:
        s.init_comm(None)
        s.load(X)
        Y = s.contract(Uc)
    ref = np.conj(X.T) @ (X @ Uc)
    assert np.linalg.norm(Y - ref) / np.linalg.norm(ref) < 1e-13


@need2
def test_e2e_two_gpus_matches_one():
    X, _ = o.spiked_samples(1500, 300, 5, 0.05, seed=21)
    S = o.sample_cov(X)
    a = sb.spEigen(S, 5, n_gpus=1)
    b = sb.spEigen(S, 5, n_gpus=2)
    b0 = sb.spEigen(S, 5, n_gpus=2, p2p_exchange=0)        # NCCL allgather of row blocks, tails replicated on all rows
    # the row-sharded tails sum their Gram partials per rank, the replicated ones over all rows at once: same result up to
    # the summation order
    nb = min(len(b["info"]["F"]), len(b0["info"]["F"]), 10)
    assert np.allclose(b["info"]["F"][:nb], b0["info"]["F"][:nb], rtol=1e-11)
    assert np.array_equal(b["vectors"] != 0, b0["vectors"] != 0)
    assert np.all(colwise_inner(b["vectors"], b0["vectors"]) > 1 - 1e-8)
    n = min(len(a["info"]["F"]), len(b["info"]["F"]), 10)
    assert np.allclose(a["info"]["F"][:n], b["info"]["F"][:n], rtol=1e-11)
    assert np.array_equal(a["vectors"] != 0, b["vectors"] != 0)
    assert np.all(colwise_inner(a["vectors"], b["vectors"]) > 1 - 1e-5)
    assert np.allclose(a["values"], b["values"], rtol=1e-12)
    c = sb.spEigen(X, 5, data=True, n_gpus=2)
    d = sb.spEigen(X, 5, data=True, n_gpus=1)
    assert np.array_equal(c["vectors"] != 0, d["vectors"] != 0)
    assert np.all(colwise_inner(c["vectors"], d["vectors"]) > 1 - 1e-5)
    assert np.allclose(c["values"], d["values"], rtol=1e-12)


@need2
def test_e2e_two_gpus_complex_cov_and_odd_sizes():
    """Row-sharded complex covariance (operand expansion waits for the peers' rows) and a size that does not divide evenly."""
    Xc, _ = o.spiked_samples(611, 700, 3, 0.1, seed=5, cplx=True)
    S = o.sample_cov(Xc)
    a = sb.spEigen(S, 3, n_gpus=1)
    b = sb.spEigen(S, 3, n_gpus=2)
    n = min(len(a["info"]["F"]), len(b["info"]["F"]), 10)
    assert np.allclose(a["info"]["F"][:n], b["info"]["F"][:n], rtol=1e-11)
    assert np.array_equal(a["vectors"] != 0, b["vectors"] != 0)
    assert np.all(colwise_inner(a["vectors"], b["vectors"]) > 1 - 1e-7)
    assert np.allclose(a["values"], b["values"], rtol=1e-12)
    Xr, _ = o.spiked_samples(1001, 250, 4, 0.05, seed=9)
    Sr = o.sample_cov(Xr)
    ar, br = sb.spEigen(Sr, 4, n_gpus=1), sb.spEigen(Sr, 4, n_gpus=2)
    n = min(len(ar["info"]["F"]), len(br["info"]["F"]), 10)
    assert np.allclose(ar["info"]["F"][:n], br["info"]["F"][:n], rtol=1e-11)
    assert np.array_equal(ar["vectors"] != 0, br["vectors"] != 0)


@need2
def test_sharded_validation_scan():
    """isSymmetric.matrix / anyNA (R/spEigen.R:53,75) on a row-sharded covariance: the transposed entries live on the peer."""
    rng = np.random.default_rng(2)
    X = rng.standard_normal((300, 500))
    S = np.cov(X, rowvar=False)
    A = S.copy(); A[400, 7] += 1e-3                       # asymmetry across the two shards
    with pytest.raises(sb.SpEigenError, match="not symmetric"): sb.spEigen(A, 3, n_gpus=2)
    r = sb.spEigen(S, 3, n_gpus=2)
    assert r["vectors"].shape == (500, 3)
    import ctypes as C
    Sn = np.asfortranarray(S.copy()); Sn[450, 450] = np.nan
    ob = sb._OutBuffers(500, 3, False)
    cfg = sb.default_cfg(n_gpus=2)
    rc = sb.lib().speigen_run_f64(Sn.ctypes.data_as(sb._dp), 500, 500, 0, 3.0, 0.5, None, None, 1e-9, C.byref(cfg), C.byref(ob.c))
    assert rc == 2


@need2
@pytest.mark.parametrize("cplx", [False, True])
def test_half_upload_row_sharded(cplx):
    """Hermitian-half upload with S row-sharded inside one process: every GPU uploads one tile of each transposed pair of its
    row block and rebuilds the others from the peers' shards over NVLink (k_mirror_cov) - same bits as the full upload."""
    import os
    rng = np.random.default_rng(11)
    m, q = 1037, 6
    A = rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0)
    S = A @ np.conj(A.T) / m + np.eye(m)
    S = (S + np.conj(S.T)) / 2                                # exactly Hermitian, positive definite
    U = rng.standard_normal((m, q)) + (1j * rng.standard_normal((m, q)) if cplx else 0)
    P = min(_ngpu(), 4)
    old = {k: os.environ.get(k) for k in ("SPEIGEN_HALF_UPLOAD", "SPEIGEN_HALF_MIN_BYTES", "SPEIGEN_HALF_BLOCK")}
    try:
        os.environ["SPEIGEN_HALF_UPLOAD"] = "0"
        with sb.Solver(m, q, is_complex=cplx, world=P, local_gpus=P) as s:
            s.init_comm(None)
            s.load(S)
            Y0 = s.contract(U)
        os.environ["SPEIGEN_HALF_UPLOAD"] = "1"
        os.environ["SPEIGEN_HALF_MIN_BYTES"] = "1"
        for block in (64, 100, 2048):
            os.environ["SPEIGEN_HALF_BLOCK"] = str(block)
            with sb.Solver(m, q, is_complex=cplx, world=P, local_gpus=P) as s:
                s.init_comm(None)
                s.load(S)
                Y1 = s.contract(U)
                s.prepare()                                   # validation on the host sums + the device scan of the rebuilt shards
            assert np.array_equal(Y0, Y1), block
            Sbad = S.copy(); Sbad[900, 17] += 1e-3            # the offending tile stays on the host for some blocks, travels for others
            with sb.Solver(m, q, is_complex=cplx, world=P, local_gpus=P) as s:
                s.init_comm(None)
                s.load(Sbad)
                with pytest.raises(sb.SpEigenError, match="not symmetric"):
                    s.prepare()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert np.linalg.norm(Y0 - S @ U) / np.linalg.norm(S @ U) < 1e-13
