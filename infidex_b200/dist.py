"""Multi-GPU plumbing for the replica (query-parallel) layout: rank-local batches, max-over-ranks timing, result exchange.

One process per GPU (torchrun), `torch.distributed` with NCCL on GPUs (gloo in the CPU tests). Queries are independent,
so there is no data-path collective; the only exchange is the per-batch all-gather of each rank's result block.
"""
import numpy as np


def rank_batch_seed(base_seed, step, rank):
    """Deterministic, distinct batch per (step, rank)."""
    return base_seed + step + 1000 * rank


def max_over_ranks(dist, values, device="cpu"):
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_results(dist, keys, device="cpu"):
    """All-gather the [nq, cap] int64 key block of every rank; returns a list indexed by rank (numpy arrays)."""
    import torch
    payload = torch.from_numpy(np.ascontiguousarray(keys)).to(device)
    out = [torch.empty_like(payload) for _ in range(dist.get_world_size())]
    dist.all_gather(out, payload)
    return [o.cpu().numpy() for o in out]
