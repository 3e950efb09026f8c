"""Candidate sharding + final selection across ranks.

The AR candidate batch is the hot path's only data-parallel axis: candidates are independent
sequences that share read-only weights, so they are sharded across ranks with NO data-path
collective.  The single exchange step is the gather of the per-candidate scores (and code
counts) for the final selection; torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests) carries it.  The winner's latents never move: the owning rank diffuses + vocodes them.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> range:
    """contiguous block partition: candidate c lives on rank c // ceil(n/world) (weights are
    replicated, so any partition works; contiguous keeps the global index = offset + local)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_select(local_scores, local_lengths, device=None):
    """all_gather (score, n_codes) of every candidate; returns (winner_global_index,
    owner_rank, local_index_on_owner, all_scores).  Ranks may hold different candidate
    counts (ragged shards are padded with -inf)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    local_scores = np.asarray(local_scores, dtype=np.float32)
    local_lengths = np.asarray(local_lengths, dtype=np.float32)
    if world == 1:
        w = int(np.argmax(local_scores))
        return w, 0, w, local_scores
    n_local = torch.tensor([len(local_scores)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    buf = torch.full((2, width), float("-inf"), dtype=torch.float32, device=device)
    buf[0, :len(local_scores)] = torch.from_numpy(local_scores)
    buf[1, :len(local_lengths)] = torch.from_numpy(local_lengths)
    allb = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    scores = np.concatenate([allb[r][0, :counts[r]].cpu().numpy() for r in range(world)])
    winner = int(np.argmax(scores))
    offs = np.cumsum([0] + counts)
    owner = int(np.searchsorted(offs, winner, side="right") - 1)
    return winner, owner, winner - int(offs[owner]), scores
