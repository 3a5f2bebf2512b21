"""CPU, world_size 2 (gloo): the particle-shard helpers of pocomc_b200.dist -- shard ranges,
rank-ordered gather of block partials and the fixed-order sum that makes sharded reductions
independent of the number of ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pocomc_b200 import dist
    assert dist.init_from_env("gloo") == (rank, world)
    assert dist.is_active()
    n, d = 1000, 5
    rng = np.random.default_rng(0)
    blocks_all = rng.normal(size=(4, d + 4))                       # ceil(1000/256) = 4 block partials of the full problem
    a, b = dist.shard_range(n, rank, world, align=256)
    mine = torch.from_numpy(blocks_all[a // 256:(b + 255) // 256].copy())
    counts = [(dist.shard_range(n, r, world, 256)[1] + 255) // 256 - dist.shard_range(n, r, world, 256)[0] // 256
              for r in range(world)]
    g = dist.gather_blocks(mine, counts)
    ok = np.array_equal(g.numpy(), blocks_all)
    # equal-count path
    eq = dist.gather_blocks(torch.full((2, 3), float(rank), dtype=torch.float64))
    ok &= eq.shape == (4, 3) and bool((eq[:2] == 0).all()) and bool((eq[2:] == 1).all())
    s = dist.allreduce_sum_det(torch.tensor([1.0 + rank, 0.1 * (rank + 1)], dtype=torch.float64))
    ok &= bool(torch.equal(s, torch.tensor([1.0, 0.1], dtype=torch.float64) + torch.tensor([2.0, 0.2], dtype=torch.float64)))
    # rows of mutated particles come back in global order (Sampler._mutate_sharded), call counters add up exactly
    rows_all = rng.normal(size=(n, 2 * d + 3))
    cnt = dist.shard_counts(n, world, 256)
    ok &= sum(cnt) == n and cnt[rank] == b - a
    back = dist.gather_blocks(torch.from_numpy(rows_all[a:b].copy()), cnt)
    ok &= np.array_equal(back.numpy(), rows_all)
    ok &= dist.allreduce_sum_int(10 ** 12 + rank) == 2 * 10 ** 12 + 1
    out[rank] = bool(ok)
    td.destroy_process_group()


def test_shard_range_covers_everything():
    from pocomc_b200 import dist
    for n, w, al in ((10000, 8, 256), (1000, 2, 256), (7, 3, 1), (256, 4, 256), (100000, 4, 256)):
        spans = [dist.shard_range(n, r, w, al) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
            assert b0 == a1 and a0 <= b0
        assert all(a % al == 0 for a, _ in spans)


def test_gather_and_deterministic_sum_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
