"""World-size-2 gloo test (CPU) of the multi-GPU host logic: utterance-batch sharding, max-over-ranks timing
and the ragged gather.  The data path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spiking_fullsubnet_b200.sharding import gather_utterances, max_over_ranks, shard_utterances


def test_shard_partition_properties():
    for n in (0, 1, 5, 32, 33, 257):
        for w in (1, 2, 3, 8):
            sl = [shard_utterances(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard_utterances(n, rank, world)
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        local = full[a:b] * 2.0  # stands for "enhance my utterances"
        out = gather_utterances(local, n)
        assert torch.equal(out, full * 2.0)
        slowest = max_over_ranks(1.0 + rank)
        assert slowest == float(world)
        # whole-job throughput as bench.py computes it: all units of all ranks / slowest rank's time
        frames = torch.tensor([float((b - a) * 501)])
        dist.all_reduce(frames)
        assert frames.item() == n * 501
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 5), nprocs=2, join=True)
