"""World-size-2 tests of the multi-GPU host logic (umap_b200/dist.py) on CPU with the gloo
backend: partitioning arithmetic, the row-sharded kNN with all-gather, the database ring with the
per-row top-k merge, and the edge-sharded gradient + all-reduce equal to the single-process result.
The compute callbacks are CPU stand-ins (the oracle); the CUDA kernels are covered by -m gpu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import umap_oracle as orc
from umap_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _knn_cpu(q, db, k, exclude_self, q_base):
    i, d = orc.knn_exact(q.numpy(), db.numpy(), k, exclude_self, self_offset=q_base)
    return torch.from_numpy(i), torch.from_numpy(d)


def _merge_cpu(ia, da, ib, db_):
    k = ia.shape[1]
    key = torch.cat([da, db_], dim=1).double() * 1e7 + torch.cat([ia, ib], dim=1).double() * 1e-3   # (dist, idx) order
    key = torch.where(torch.cat([ia, ib], dim=1) < 0, torch.full_like(key, float("inf")), key)
    order = torch.argsort(key, dim=1, stable=True)[:, :k]
    return torch.gather(torch.cat([ia, ib], 1), 1, order), torch.gather(torch.cat([da, db_], 1), 1, order)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        n, d, k = 300, 12, 7
        x = torch.from_numpy((rng.standard_normal((n, d)) + 3 * rng.integers(0, 3, (n, 1))).astype(np.float32))
        # 1. row-sharded kNN + all-gather
        idx, dd = D.knn_sharded_rows(x, k, True, _knn_cpu)
        ri, rd = orc.knn_exact(x.numpy(), x.numpy(), k, True)
        assert np.array_equal(idx.numpy(), ri) and np.array_equal(dd.numpy().view(np.uint32), rd.view(np.uint32))
        # 2. database ring with running merge: rank owns row_block
        lo, hi = D.row_block(n, rank, world)
        bi, bd = D.ring_knn(x[lo:hi].contiguous(), n, k, True, _knn_cpu, _merge_cpu)
        assert np.array_equal(bi.numpy(), ri[lo:hi]), "ring kNN differs"
        assert np.array_equal(bd.numpy().view(np.uint32), rd[lo:hi].view(np.uint32))
        # 3. edge-sharded gradient: batches of this rank only, global per-batch normalisation, all-reduce
        n_b, bs, num_rep, dim = 5, 64, 4, 4
        y = rng.standard_normal((n, dim))
        rows = np.repeat(np.arange(n), k)
        cols = ri.reshape(-1).astype(np.int64)
        keep = rng.random(rows.shape[0]) < 0.4
        negs = rng.integers(0, n, (int(keep.sum()), num_rep))
        ii, jj = rows[keep], cols[keep]
        batch = ii // bs

        def grad_of(sel):
            g = np.zeros_like(y)
            for b in np.unique(batch[sel]):
                m = sel & (batch == b)
                _, ga = orc.umap_attr_grad(y[ii[m]], y[jj[m]], 1.577, 0.8951)       # mean over the batch's kept edges
                np.add.at(g, ii[m], ga / n_b)
                np.add.at(g, jj[m], -ga / n_b)
                ir = np.repeat(ii[m], num_rep)
                _, gr = orc.umap_rep_grad(y[ir], y[negs[m].reshape(-1)], 1.577, 0.8951)
                np.add.at(g, ir, gr / n_b)
                np.add.at(g, negs[m].reshape(-1), -gr / n_b)
            return g

        b_lo, b_hi = D.batch_range(n_b, rank, world)
        mine = (batch >= b_lo) & (batch < b_hi)
        g_local = torch.from_numpy(grad_of(mine))
        D.all_reduce_sum(g_local)
        g_full = grad_of(np.ones_like(mine))
        assert np.allclose(g_local.numpy(), g_full, rtol=1e-12, atol=1e-15)
        assert D.same_on_all_ranks(100 + rank) == 100
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_partition_arithmetic():
    for n in (1, 127, 128, 129, 31783, 158915):
        for w in (1, 2, 4, 8):
            blocks = [D.row_block(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(lo % 128 == 0 for lo, hi in blocks if hi > lo)
            assert all(hi - lo <= D.block_size(n, w) for lo, hi in blocks)
    for nb in (1, 5, 621):
        for w in (1, 2, 8):
            rs = [D.batch_range(nb, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == nb and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    # shards balanced by weight: contiguous, covering, and each rank's weight within one batch of total / w
    rng = np.random.default_rng(4)
    for nb in (1, 7, 621):
        wts = rng.random(nb) ** 3 * 100.0                       # strongly non-uniform batches
        cum = np.cumsum(wts).tolist()
        for w in (1, 2, 4, 8):
            rs = [D.balanced_batch_range(cum, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == nb and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(lo <= hi for lo, hi in rs)
            if nb >= 8 * w:
                share = [wts[lo:hi].sum() for lo, hi in rs]
                assert max(share) <= wts.sum() / w + wts.max() + 1e-9
    assert D.balanced_batch_range([0.0, 0.0, 0.0], 1, 2) == D.batch_range(3, 1, 2)


@pytest.mark.timeout(300)
def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
