"""Launched under torchrun (one rank per GPU): the item-sharded sampler must reproduce the single-GPU chain.
theta draws identical (grid points); beta / f of each shard equal the corresponding columns of the unsharded run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist

import gpirt_b200.sampler as G
from gpirt_b200 import ResponseMatrix
from conftest import make_problem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid_t = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid_t = torch.tensor(list(G.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid_t, 0)
uid = bytes(uid_t.cpu().tolist())
n, m, S, B = int(os.environ.get("SHARD_N", "300")), int(os.environ.get("SHARD_M", "90")), 3, 2
missing = float(os.environ.get("SHARD_MISSING", "0.05"))
p = make_problem(n, m, seed=5, missing=missing)
from gpirt_b200.sharding import item_block
j0, j1 = item_block(m, rank, world)
got = G.gpirtMCMC(ResponseMatrix(p["y"][:, j0:j1]), S, B, beta_prior_means=p["pm"][:, j0:j1], beta_prior_sds=p["psd"][:, j0:j1],
                  beta_proposal_sds=p["pstep"][:, j0:j1], theta_init=p["theta"], seed=99, device=local,
                  shard=(rank, world, m, j0, uid))
ok = True
if rank == 0:
    full = G.gpirtMCMC(ResponseMatrix(p["y"]), S, B, beta_prior_means=p["pm"], beta_prior_sds=p["psd"],
                       beta_proposal_sds=p["pstep"], theta_init=p["theta"], seed=99, device=local)
    th_same = np.array_equal(got["theta"], full["theta"])
    db = np.max(np.abs(got["beta"] - full["beta"][:, j0:j1]))
    df = np.max(np.abs(got["f"] - full["f"][:, j0:j1]))
    di = np.max(np.abs(got["IRFs"] - full["IRFs"][:, j0:j1]))
    ok = th_same and db <= 1e-9 and df <= 1e-7 and di <= 1e-7
    print("SHARD_CHECK world=%d missing=%.2f theta_identical=%s max|dbeta|=%.2e max|df|=%.2e max|dIRF|=%.2e -> %s" %
          (world, missing, th_same, db, df, di, "OK" if ok else "FAIL"), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
