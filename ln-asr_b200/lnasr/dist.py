"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The acoustic path shards by UTTERANCE: decoding needs no data-path collective; Baum-Welch
needs one all-reduce (SUM, float64) of the packed accumulator buffer per EM iteration, after which every rank
runs the same M-step kernel on identical numbers, so replicas stay bit-identical.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

__all__ = ["shard_utterances", "make_reduce_fn", "world"]


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_utterances(lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` owns: longest-first greedy assignment to the least loaded rank, so
    frames per rank differ by at most one utterance.  Deterministic; every rank computes the same partition."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lengths[i]
    return [int(i) for i in np.nonzero(owner == rank)[0]]


def make_reduce_fn(group=None):
    """In-place SUM all-reduce of a tensor over `group`; identity when there is a single rank."""
    import torch.distributed as dist

    def reduce_fn(t):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t
    return reduce_fn
