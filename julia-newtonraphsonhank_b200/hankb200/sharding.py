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


def period_round_robin(n, world, rank, n_endog=4):
    """Balanced alternative to column_partition for Jacobian builds: the periods are dealt round-robin, rank r
    takes all n_endog columns of the periods t with t % world == r (1-based ascending column list).  A unit seed at
    period s costs a backward sweep of s periods, so contiguous blocks would leave every long sweep on the last
    rank; dealt this way every rank holds the same mix of seed horizons."""
    P = n // n_endog
    return np.array([n_endog * t + v + 1 for t in range(rank, P, world) for v in range(n_endog)], dtype=np.int32)


def round_robin_permutation(n, world, n_endog=4):
    """perm such that J_full[:, perm] = concatenation over ranks of the period_round_robin blocks."""
    return np.concatenate([period_round_robin(n, world, r, n_endog) - 1 for r in range(world)])


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


def jacobian_distributed(blk, n, world, rank, group=None, balanced=False):
    """Full (n, n) Jacobian at blk's current linearisation: each rank builds its column block with
    `blk.jacobian_columns(begin, end)` (contiguous blocks) or, with `balanced`, `blk.jacobian_column_list(cols)`
    for its round-robin periods, and the blocks are gathered (and put back in column order)."""
    if balanced:
        cols = period_round_robin(n, world, rank)
        local = blk.jacobian_column_list(cols) if len(cols) else np.zeros((n, 0))
        if world == 1:
            return local
        J = np.empty((n, n))
        J[:, round_robin_permutation(n, world)] = gather_columns(local, world, rank, group)
        return J
    b, e = column_partition(n, world)[rank]
    local = blk.jacobian_columns(b, e) if e > b else np.zeros((n, 0))
    if world == 1:
        return local
    return gather_columns(local, world, rank, group)
