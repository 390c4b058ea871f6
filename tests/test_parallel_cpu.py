"""Host-side multi-rank logic on CPU: sharding arithmetic and the one all-gather,
world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magi_v2_b200 import parallel


def test_shard_ranges_cover_and_balance():
    for n in (1, 7, 20, 4096, 4099):
        for world in (1, 2, 3, 8):
            rs = [parallel.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = parallel.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
            assert [parallel.chain_id0(n, 8, r, world) for r in range(world)] == [8 * lo for lo, _ in rs]
    with pytest.raises(ValueError):
        parallel.shard_range(4, 4, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = parallel.shard_range(B, rank, world)
        # samples [n_iter, B_local, R, P]: value encodes (iteration, global dataset, chain, parameter)
        it, R, P = 3, 2, 3
        g = torch.arange(lo, hi, dtype=torch.float64)
        local = (torch.arange(it, dtype=torch.float64)[:, None, None, None] * 1e6 + g[None, :, None, None] * 1e3
                 + torch.arange(R, dtype=torch.float64)[None, None, :, None] * 10
                 + torch.arange(P, dtype=torch.float64)[None, None, None, :])
        out = parallel.gather_samples(local, dataset_dim=1, sizes=parallel.shard_sizes(B, world))
        q.put((rank, out.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 8), (2, 7), (3, 10)])
def test_gather_samples_gloo(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    it, R, P = 3, 2, 3
    g = np.arange(B, dtype=np.float64)
    want = (np.arange(it)[:, None, None, None] * 1e6 + g[None, :, None, None] * 1e3
            + np.arange(R)[None, None, :, None] * 10 + np.arange(P)[None, None, None, :])
    for _, out in res:
        assert out.shape == (it, B, R, P)
        assert np.array_equal(out, want)


def test_gather_is_identity_without_process_group():
    x = torch.arange(12, dtype=torch.float64).reshape(2, 3, 2)
    assert parallel.gather_samples(x) is x
