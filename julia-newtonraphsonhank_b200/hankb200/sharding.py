"""Lane / column sharding for multi-GPU Jacobian and JVP batches.

Only tangent lanes shard: every rank recomputes the primal and carries its own block of columns; the
blocks are all-gathered.  The data path on GPUs is `hank_allgather_columns_dev` (NCCL); this module
holds the host-side partition / reassembly logic and a torch.distributed gather used on CPU (gloo) in
tests and by `jacobian_distributed`.
"""
import numpy as np


def column_partition(n, world):
    """Contiguous 1-based half-open column ranges [begin, end) per rank, sizes differing by <= 1.
    Variable-fastest ordering (Y, KS, r, w per period) keeps the r/w columns that need household sweeps
    evenly spread."""
    base, rem = divmod(n, world)
    out, b = [], 1
    for r in range(world):
        k = base + (1 if r < rem else 0)
        out.append((b, b + k))
        b += k
    return out


def lane_slice(K, world, rank):
    """Lanes [lo, hi) of a K-lane JVP batch owned by `rank`."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_columns(local, world, rank, group=None):
    """All-gather (n, k_r) column blocks of unequal width into (n, sum k_r) on every rank."""
    import torch
    import torch.distributed as dist
    n = local.shape[0]
    widths = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(widths, torch.tensor([local.shape[1]], dtype=torch.int64), group=group)
    widths = [int(w.item()) for w in widths]
    kmax = max(widths)
    pad = np.zeros((kmax, n))
    pad[: local.shape[1]] = local.T          # column-major blocks, like the device layout
    bufs = [torch.zeros((kmax, n), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad), group=group)
    return np.concatenate([bufs[r].numpy()[: widths[r]] for r in range(world)], axis=0).T


def jacobian_distributed(blk, n, world, rank, group=None):
    """Full (n, n) Jacobian at blk's current linearisation: each rank builds its column block with
    `blk.jacobian_columns(begin, end)` and the blocks are gathered."""
    b, e = column_partition(n, world)[rank]
    local = blk.jacobian_columns(b, e) if e > b else np.zeros((n, 0))
    if world == 1:
        return local
    return gather_columns(local, world, rank, group)
