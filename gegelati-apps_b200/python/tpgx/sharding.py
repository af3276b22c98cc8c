"""Root sharding across ranks (one process per GPU) and the gather of per-root score records.

The reference parallelises evaluateAllRoots over roots with threads and merges results by job index
([UP] ParallelLearningAgent::evaluateAllRootsInParallel).  Here rank r of `world` evaluates a contiguous
slice of the job-index order and ONE all-gather of fixed-size records makes every score visible
everywhere; nothing else on the path communicates."""
import numpy as np


def shard_range(rank, world, nb_roots):
    """Contiguous slice [begin, end) of the root list for `rank`; slices differ by at most one root."""
    if not (0 <= rank < world) or nb_roots < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(nb_roots, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slot_shard(rank, world, nb_roots):
    """The library's own convention (tpgx_comm_shard): slot = ceil(nb_roots / world) roots per rank, rank r evaluates
    [r * slot, min((r + 1) * slot, nb_roots)) — the gathered buffer is then indexed by the global root number directly."""
    if not (0 <= rank < world) or nb_roots < 0:
        raise ValueError("bad shard request")
    slot = -(-nb_roots // world)
    return min(nb_roots, rank * slot), min(nb_roots, (rank + 1) * slot)


def gather_records(local_records, world, nb_roots, dist=None):
    """All-gathers [n_local, 1 + C] float64 records into [nb_roots, 1 + C] in job-index order.
    Works on torch tensors of any device (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch
    if world == 1:
        return local_records
    if dist is None:
        import torch.distributed as dist
    width = local_records.shape[1]
    per = -(-nb_roots // world)  # shards padded to the largest one so a single fixed-size collective suffices
    padded = torch.zeros(per, width, dtype=local_records.dtype, device=local_records.device)
    padded[: local_records.shape[0]] = local_records
    out = torch.empty(world * per, width, dtype=local_records.dtype, device=local_records.device)
    dist.all_gather_into_tensor(out, padded)
    rows = []
    for r in range(world):
        b, e = shard_range(r, world, nb_roots)
        rows.append(out[r * per: r * per + (e - b)])
    return torch.cat(rows, 0)


def records_from(score, class_f1):
    """[n, 1 + C] record layout of tpgx_evaluate_classification_device from host result arrays."""
    return np.concatenate([np.asarray(score)[:, None], np.asarray(class_f1)], axis=1)
