"""Utterance sharding across GPUs: nothing couples utterances inside the solve, so a batch is
dealt to one replica per GPU by cost with the longest-processing-time rule; no collective on the data path."""
from __future__ import annotations

from typing import List, Sequence


def utterance_cost(length: int, channels: int) -> float:
    c = channels
    return length * (274.0 * c * c + 1800.0 * c) + 24.0 * c * length * length


def shard_utterances(lengths: Sequence[int], world_size: int, channels: int = 384) -> List[List[int]]:
    """Returns, per rank, the indices (into ``lengths``) of its utterances; deterministic for a given input."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-utterance_cost(lengths[i], channels), i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += utterance_cost(lengths[i], channels)
    return [sorted(s) for s in shards]


def gather_outputs(per_rank_outputs, shards, batch: int):
    """Inverse of shard_utterances on the host: per_rank_outputs[r][j] is the mel of utterance shards[r][j]."""
    out = [None] * batch
    for r, idxs in enumerate(shards):
        for j, i in enumerate(idxs):
            out[i] = per_rank_outputs[r][j]
    return out
