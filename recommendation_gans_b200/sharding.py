"""Multi-GPU plumbing for the evaluation path (one process per GPU, torch.distributed).

Evaluation shards naturally: users are independent, the item table is replicated.  Each rank scores a
contiguous block of the evaluated users and the per-k hit sums are all-reduced.  Training of
the replicated-table configs does not shard without changing the reference's math ("replicas only").
These helpers hold the host-side logic so it can be tested on CPU with the gloo backend."""
import numpy as np
import torch


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units for `rank`; blocks differ in size by at most one."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_precision_recall(hits, ntargets, ks, dist=None, device=None):
    """Mean precision@k / recall@k over ALL ranks' users from this rank's per-user hit counts.

    hits: int array [n_local_users, len(ks)], ntargets: int array [n_local_users].
    Returns (precision[len(ks)], recall[len(ks)], n_users_total): the same numbers a single process would
    compute over the union of the shards (sums are accumulated in float64; the reduction order over ranks is fixed
    by all_reduce, so every rank gets identical values)."""
    ks = np.asarray(ks, dtype=np.float64)
    hits = np.asarray(hits, dtype=np.float64).reshape(-1, len(ks))
    ntargets = np.asarray(ntargets, dtype=np.float64).reshape(-1)
    sums = np.concatenate([(hits / ks[None, :]).sum(0), (hits / ntargets[:, None]).sum(0) if len(hits) else np.zeros(len(ks)),
                           [float(len(hits))]])
    t = torch.from_numpy(sums)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    tot = t.numpy()
    n = tot[-1]
    nk = len(ks)
    return tot[:nk] / n, tot[nk:2 * nk] / n, int(n)
