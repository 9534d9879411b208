"""world_size-2 gloo test of the N>1 exchange logic on CPU: the shard partition exported by the C ABI,
allgather of row blocks (covariance) and rank-ordered sum of partial panels (data matrix) reproduce the
unsharded contraction.  The local products are done by the oracle (NumPy) — test infrastructure only."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import sparseeigen_b200 as sb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                     # same stream on every rank
        m, n, q = 37, 53, 3
        X = rng.standard_normal((n, m))
        S = X.T @ X / (n - 1)
        U = rng.standard_normal((m, q))
        # --- covariance: row-sharded S, allgather of equal row blocks -----------------------------
        lo, hi = sb.shard_range(m, world, rank)
        per = -(-m // world)
        pitch = sb.panel_pitch(q)
        block = np.zeros((per, pitch))
        block[:hi - lo, :q] = S[:, lo:hi].T @ U            # the column slab of S is the row block (S = S^T)
        gathered = [torch.zeros(per, pitch, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(block))
        Y = torch.cat(gathered)[:m, :q].numpy()
        ok_cov = np.allclose(Y, S @ U, rtol=1e-13, atol=1e-13)
        # --- data matrix: sample-sharded X, allgather of partial panels + fixed-order sum ----------
        lo, hi = sb.shard_range(n, world, rank)
        part = X[lo:hi].T @ (X[lo:hi] @ U)
        parts = [torch.zeros(m, q, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(part))
        G = np.zeros((m, q))
        for r in range(world):                             # rank order => identical bits on every rank
            G = G + parts[r].numpy()
        ok_data = np.allclose(G, X.T @ (X @ U), rtol=1e-12, atol=1e-12)
        # --- cov(X) formed shard by shard (speigen_solver_load_cov_of_data): rank r computes the column slab
        #     S[:, lo:hi] = Xc^T Xc[:, lo:hi] / (n - 1) of its rows; together the slabs are cov(X) ------------------------
        lo, hi = sb.shard_range(m, world, rank)
        Xc = X - X.mean(axis=0, keepdims=True)
        slab = np.zeros((m, per))
        slab[:, :hi - lo] = Xc.T @ Xc[:, lo:hi] / (n - 1)
        slabs = [torch.zeros(m, per, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(slabs, torch.from_numpy(slab))
        Sfull = torch.cat(slabs, dim=1)[:, :m].numpy()
        ok_covx = np.allclose(Sfull, np.cov(X, rowvar=False), rtol=1e-13, atol=1e-13)
        ok_cov = ok_cov and ok_covx
        bits = torch.from_numpy(G.copy())
        ref = bits.clone()
        dist.broadcast(ref, src=0)
        ok_bits = bool(torch.equal(bits, ref))
        ret[rank] = (ok_cov, ok_data, ok_bits)
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == (True, True, True), (r, ret[r])
