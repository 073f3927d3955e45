"""Host-side item sharding for multi-GPU runs (one process per GPU).

Items (columns of y; f, beta, f*, IRFs with them) are split into contiguous blocks, one per rank; theta, K, its Cholesky
factor and the grid stay replicated.  The only exchange per sweep is the sum over ranks of the per-respondent
log-posterior partials of the theta step (reference src/draw-theta.cpp:18 sums over ALL items), done inside the CUDA
library with one NCCL all-reduce.  Random variates are addressed by GLOBAL item index (item_offset + local j), so a
sharded run reproduces the single-GPU chain."""
import numpy as np


def item_block(m, rank, world):
    """[j0, j1) of the contiguous item block owned by `rank`: balanced blocks (sizes differ by at most one), so no rank is
    left without items as long as m >= world (the library rejects m_global < world_size on every rank)."""
    return (rank * m) // world, ((rank + 1) * m) // world


def share_unique_id(dist, rank, make_uid, device=None):
    """Rank 0 creates a 128-byte NCCL unique id with make_uid(); torch.distributed (any backend) carries it to all ranks.
    One id per communicator, i.e. per sampler / gpirtMCMC call."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t = torch.tensor(list(make_uid()), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def gather_items(dist, local, m, rank, world, axis=1):
    """All-gather item-sharded host arrays (e.g. beta (2, m_loc, S+1) or IRFs (1001, m_loc)) into the full array."""
    import torch
    per = (m + world - 1) // world
    pad_shape = list(local.shape)
    pad_shape[axis] = per
    buf = np.zeros(pad_shape, dtype=np.float64)
    sl = [slice(None)] * local.ndim
    sl[axis] = slice(0, local.shape[axis])
    buf[tuple(sl)] = local
    out = [torch.zeros(pad_shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(out, torch.from_numpy(buf))
    parts = []
    for r in range(world):
        j0, j1 = item_block(m, r, world)
        sl[axis] = slice(0, j1 - j0)
        parts.append(out[r].numpy()[tuple(sl)])
    return np.concatenate(parts, axis=axis)
