"""world_size-2 gloo test of the utterance-sharded decode plumbing: each rank takes its shard,
'decodes' it (here: a deterministic stand-in, no GPU), and the host gathers by original index."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import matcha_tts_24k_b200 as P
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [50, 7, 33, 12, 90, 41, 5]
    shards = P.shard_utterances(lengths, world, channels=64)
    mine = shards[rank]
    outs = [torch.full((3, lengths[i]), float(i)) for i in mine]  # stand-in for the per-utterance mel
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)
    t = torch.tensor([float(sum(lengths[i] for i in mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the bench's max-over-ranks reduction
    if rank == 0:
        full = P.gather_outputs(gathered, shards, len(lengths))
        ok = all(full[i].shape == (3, lengths[i]) and bool((full[i] == i).all()) for i in range(len(lengths)))
        ret["ok"] = ok and float(t) >= sum(lengths) / world
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret.get("ok") is True
