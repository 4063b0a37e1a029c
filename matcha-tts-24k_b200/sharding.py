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


class ShardedCFM:
    """The decode of one batch across every GPU of the box, in ONE process (BASELINE config 3).

    One ``CFM`` replica (library handle + full weight copy + its own plans / CUDA graphs) per device.  ``solve_host`` deals the
    utterances to the replicas by cost (longest processing time first), plans each replica for its own lengths (the batch-wide
    ``T`` stays: GroupNorm and the additive attention mask see the padded length), enqueues on every GPU - per-utterance H2D from
    the caller's host tensors, the decode, per-utterance D2H into the caller's result tensor at the utterance's ORIGINAL index
    (``cfm_solve_host_indexed``) - and only then waits for them.  No collective, no NCCL: utterances never interact.
    The enqueue calls run on one host thread per GPU (ctypes releases the GIL), so a replica's first decode of a new shape,
    which launches ~1200 kernels directly, does not delay the other GPUs.
    """

    def __init__(self, in_channels, out_channel, cfm_params, decoder_params, devices=None, precision=None, flags: int = 0):
        import torch
        from .cfm import CFM
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("ShardedCFM needs at least one CUDA device: there is no CPU path")
        self.devices = [torch.device("cuda", int(d)) if not isinstance(d, torch.device) else d for d in devices]
        self.replicas = [CFM(in_channels, out_channel, cfm_params, decoder_params, precision=precision, flags=flags).eval()
                         for _ in self.devices]
        for r, d in zip(self.replicas, self.devices):
            r.to(d)
        self.solver = cfm_params.solver
        self.channels = self.replicas[0]._weights.cfg.channels
        self._pool = None
        self.last_shards = None

    @property
    def estimator(self):
        return self.replicas[0].estimator

    def load_estimator_state_dict(self, state_dict):
        """Fills every replica's estimator (keys as below ``decoder.estimator.`` in the reference checkpoint)."""
        for r in self.replicas:
            r.estimator.load_state_dict(state_dict)

    def close(self):
        for r in self.replicas:
            r.close()
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None

    def solve_host(self, x, t_span, mu, lengths, spks=None, out=None):
        """x (initial state / injected noise), mu: host tensors (B, n_feats, T), ideally pinned; returns the mel (B, n_feats, T)
        on the host, padded frames equal to x there (reference flow_matching.py:60-63 for every utterance)."""
        import ctypes as C
        import torch
        from concurrent.futures import ThreadPoolExecutor
        from . import _native as N
        if x.is_cuda or mu.is_cuda:
            raise ValueError("ShardedCFM.solve_host takes host tensors")
        mu_, x_ = mu.detach().float().contiguous(), x.detach().float().contiguous()
        r0 = self.replicas[0]
        B, F, T = r0._check_shapes(mu_, x_)
        lengths = r0._check_lengths(lengths, B, T)
        spks_ = r0._check_spks(spks, B, prep=False)
        if out is None:
            out = torch.empty_like(mu_)
        elif out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != tuple(mu_.shape) or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 host tensor of mu's shape")
        ts = [float(v) for v in torch.as_tensor(t_span, dtype=torch.float32).tolist()]
        shards = [s for s in shard_utterances(lengths, len(self.replicas), self.channels)]
        self.last_shards = shards
        work = [(r, d, s) for r, d, s in zip(self.replicas, self.devices, shards) if s]

        def enqueue(item):
            r, d, idx = item
            r.solver = self.solver
            lib, handle = r._ensure(d, [lengths[i] for i in idx], T, ts, self.solver)
            arr = (C.c_int32 * len(idx))(*idx)
            N.check(lib, handle, lib.cfm_solve_host_indexed(handle, mu_.data_ptr(), x_.data_ptr(),
                                                            spks_.data_ptr() if spks_ is not None else None, out.data_ptr(), arr, B))
            return lib, handle

        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=len(self.replicas))
        started = list(self._pool.map(enqueue, work))   # every GPU has its work before the host waits for any
        for lib, handle in started:
            N.check(lib, handle, lib.cfm_synchronize(handle))
        return out
