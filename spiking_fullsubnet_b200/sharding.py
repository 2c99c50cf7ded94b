"""Multi-GPU host logic of the hot path: the path shards by UTTERANCE BATCH only -- rows of
different utterances are independent recurrences, while the sub-bands of one utterance share its full-band
output and therefore stay on one GPU.  One process per GPU; inference needs no data-path collective.  Only
`torch.distributed` plumbing lives here (works on gloo/CPU for tests and nccl on B200)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_utterances(num_utterances, rank, world_size):
    """Contiguous, balanced [start, stop) slice of the utterance batch owned by `rank` (first ranks get the
    remainder).  Every utterance belongs to exactly one rank; empty slices are legal."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(num_utterances, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (device-time reporting: the slowest rank defines the step)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_utterances(local, num_utterances, dim=0):
    """All-gather per-rank results computed on `shard_utterances` slices back into batch order (used when a
    caller wants every enhanced utterance on every rank, e.g. validation metrics).  Ragged shards allowed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_utterances(num_utterances, r, world) for r in range(world)]
    longest = max(b - a for a, b in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = longest
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded)
    return torch.cat([o.narrow(dim, 0, b - a) for o, (a, b) in zip(outs, sizes)], dim=dim)
