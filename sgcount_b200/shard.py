"""Partitioning of the count path over ranks (one process per GPU) and the one exchange step.

The reference's only fan-out is rayon over samples (count.rs:117-136); its per-sample
`Counter`s are collected into a Vec (count.rs:136).  Here a rank owns whole samples when there
are at least as many samples as ranks, and a contiguous read range of a sample otherwise; the
per-guide count vectors `u64[n_guides + 2]` (counts, total_reads, matched_reads) are the only
state that crosses ranks and are summed with one all-reduce per sample that spans ranks.
Offsets are detected once per sample (on the sample's first `subsample` records) by the rank
that owns the sample's first shard and broadcast, never per shard.

Backend-agnostic: tensors live wherever the process group expects them (CUDA for nccl, CPU
for gloo, which is what the CPU tests use).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    sample: int
    first_read: int
    n_reads: int
    rank: int


def plan_shards(reads_per_sample: Sequence[int], world: int, align: int = 256) -> List[Shard]:
    """Every read of every sample appears in exactly one shard.

    n_samples >= world: samples are dealt to ranks whole, largest first onto the least loaded
    rank (no exchange needed).  Otherwise every sample is cut into `world` contiguous read
    ranges whose boundaries are multiples of `align` reads (keeps fixed-stride shards 16-byte
    aligned for the streaming kernel)."""
    n = len(reads_per_sample)
    shards: List[Shard] = []
    if world <= 1:
        return [Shard(s, 0, int(r), 0) for s, r in enumerate(reads_per_sample)]
    if n >= world:
        load = [0] * world
        for s in sorted(range(n), key=lambda i: (-reads_per_sample[i], i)):
            r = min(range(world), key=lambda i: (load[i], i))
            load[r] += int(reads_per_sample[s])
            shards.append(Shard(s, 0, int(reads_per_sample[s]), r))
        return sorted(shards, key=lambda x: (x.sample, x.first_read))
    for s, total in enumerate(reads_per_sample):
        total = int(total)
        per = -(-total // world)
        per = -(-per // align) * align
        for r in range(world):
            a = min(total, r * per)
            b = min(total, (r + 1) * per)
            if b > a or (r == 0 and total == 0):
                shards.append(Shard(s, a, b - a, r))
    return shards


def samples_spanning_ranks(shards: Sequence[Shard]) -> Dict[int, List[int]]:
    """sample -> sorted ranks that hold a shard of it, for samples held by more than one rank"""
    owners: Dict[int, set] = {}
    for sh in shards:
        owners.setdefault(sh.sample, set()).add(sh.rank)
    return {s: sorted(r) for s, r in owners.items() if len(r) > 1}


def owner_of(shards: Sequence[Shard], sample: int) -> int:
    """rank that holds the sample's first reads: detects the offset and writes the column"""
    return min((sh for sh in shards if sh.sample == sample), key=lambda x: x.first_read).rank


def reduce_counts(state: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the ranks' `int64[n_guides + 2]` state vectors, in place (the NCCL / gloo
    all-reduce standing for count.rs:136's collect).  uint64 counts travel as int64: the sum
    is the same bit pattern."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(state, op=dist.ReduceOp.SUM, group=group)
    return state


def broadcast_offset(offset, src: int, device=None, group=None):
    """(reverse, index) decided by `src`, handed to every rank"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return offset
    t = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank(group) == src:
        t[0], t[1] = int(offset[0]), int(offset[1])
    dist.broadcast(t, src=src, group=group)
    return bool(t[0].item()), int(t[1].item())


def count_samples(shards: Sequence[Shard], n_samples: int, n_guides: int, rank: int, count_shard,
                  device=None, group=None) -> torch.Tensor:
    """Runs `count_shard(shard) -> int64[n_guides + 2]` (tensor on `device`) for this rank's
    shards and returns the full `int64[n_samples, n_guides + 2]` table on every rank.

    Columns of samples held by one rank need no arithmetic, only delivery; summing a matrix
    whose other rows are zero does both with ONE collective over `n_samples x (n_guides+2)`
    words (<= 12.8 MB at 8 x 200 k guides), which is cheaper than per-sample calls."""
    table = torch.zeros((n_samples, n_guides + 2), dtype=torch.int64, device=device)
    for sh in shards:
        if sh.rank == rank:
            table[sh.sample] += count_shard(sh)
    return reduce_counts(table, group)
