"""CPU, world_size 2, gloo: the multi-GPU host logic (batch sharding, global-noise slicing, the single all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_layout_b200 import parallel


def test_shard_range_partitions_exactly():
    for gb in (0, 1, 7, 8, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [parallel.shard_range(gb, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (gb, 1, 4, 8)
        x_T, noise = parallel.global_noise(shape, seed=1000, n_steps=3)
        xl = parallel.local_slice(x_T, rank, world)
        nl = parallel.local_slice(noise, rank, world, batch_dim=1)
        lo, hi = parallel.shard_range(gb, rank, world)
        assert xl.shape[0] == hi - lo and nl.shape[:2] == (3, hi - lo)
        # stand-in for the per-rank sampling: any per-sample function
        local_img = xl * 2 + nl.sum(0)
        full = parallel.all_gather_batch(local_img, gb)
        ref = x_T * 2 + noise.sum(0)
        q.put((rank, bool(torch.equal(full, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gb", [8, 5])
def test_two_rank_gather_reproduces_single_rank(gb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
