"""File-shard bookkeeping for the multi-GPU run: independent capture files,
file i -> rank i mod world_size, no exchange step on the data path.  torch.distributed is used
only to agree on the slowest rank's time and to sum the counters."""
import torch
import torch.distributed as dist


def files_for_rank(n_files, rank, world):
    """Round-robin shard: the global file indices this rank owns, in order."""
    return list(range(rank, n_files, world))


def owner_of(file_index, world):
    return file_index % world


def reduce_report(samples, packages, events, elapsed_ms, device=None):
    """-> (sum samples, sum packages, sum events, max elapsed_ms) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return samples, packages, events, elapsed_ms
    t = torch.tensor([float(samples), float(packages), float(events)], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    s, p, e = t.tolist()
    return int(s), int(p), int(e), m.item()
