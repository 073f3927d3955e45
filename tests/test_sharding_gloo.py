"""World-size-2 CPU test (gloo) of the host-side multi-GPU logic: item partition, unique-id distribution, gathers, and
the identity the sharded theta step relies on — per-respondent log-likelihood partials summed over item shards equal
the full-matrix value (checked with the oracle as the arithmetic stand-in; no CUDA involved)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import make_problem
    from gpirt_b200.sharding import gather_items, item_block, share_unique_id
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, m = 25, 13
        p = make_problem(n, m, seed=3, missing=0.1)
        j0, j1 = item_block(m, rank, world)
        # 1. unique id travels from rank 0
        uid = share_unique_id(dist, rank, lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2. partial log-likelihood sums over the local items, all-reduced, equal the full computation
        ts, prior = O.grid()
        fstar = np.asfortranarray(np.random.RandomState(1).randn(1001, m))
        _, _, lp_loc = O.draw_theta(ts, p["y"][:, j0:j1], np.zeros(1001), fstar[:, j0:j1], O.Rng.keyed(1), mode=1)
        t = torch.from_numpy(np.ascontiguousarray(lp_loc))
        dist.all_reduce(t)
        _, idx_full, lp_full = O.draw_theta(ts, p["y"], np.zeros(1001), fstar, O.Rng.keyed(1), mode=1)
        assert np.max(np.abs(t.numpy() - lp_full)) <= 1e-11 * np.abs(lp_full).max()
        # 3. ESS draws addressed by GLOBAL item index: a shard reproduces its columns of the unsharded step
        L = O.build_cholS(p["theta"])
        f = np.asfortranarray(L @ np.random.RandomState(2).randn(n, m))
        mu = np.zeros((n, m))
        full = np.stack([O.ess(f[:, j], p["y"][:, j], L, mu[:, j], j, _rng(O))[0] for j in range(m)], axis=1)
        mine = np.stack([O.ess(f[:, j], p["y"][:, j], L, mu[:, j], j, _rng(O))[0] for j in range(j0, j1)], axis=1)
        assert np.array_equal(mine, full[:, j0:j1])
        # 4. gather of item-sharded results
        beta_loc = np.arange(2 * (j1 - j0) * 3, dtype=np.float64).reshape(2, j1 - j0, 3) + 1000 * rank
        allb = gather_items(dist, beta_loc, m, rank, world, axis=1)
        assert allb.shape == (2, m, 3) and np.array_equal(allb[:, j0:j1], beta_loc)
        open(os.path.join(tmpdir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _rng(O):
    r = O.Rng.keyed(77)
    r.set_sweep(4)
    return r


def test_item_block_tiles_the_items():
    from gpirt_b200.sharding import item_block
    for m in (1, 7, 10, 10000):
        for world in (1, 2, 3, 4, 8):
            blocks = [item_block(m, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
