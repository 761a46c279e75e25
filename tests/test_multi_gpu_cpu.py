"""World-size-2 gloo test of the utterance sharding / result gathering used by bench.py --gpus N (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sepreformer_b200.sharding import gather_utterance_values, shard_bounds


def test_shard_bounds_cover_everything_once():
    for total in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, world, rank)
        # per-utterance "metric": a function of the global utterance index, two speakers
        local = torch.stack([torch.arange(lo, hi, dtype=torch.float32) * 10 + s for s in (0, 1)], dim=1)
        full = gather_utterance_values(local, total)
        want = torch.stack([torch.arange(total, dtype=torch.float32) * 10 + s for s in (0, 1)], dim=1)
        ret[rank] = bool(torch.equal(full, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_reconstructs_global_order_two_ranks(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}
