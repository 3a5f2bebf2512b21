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
    # persistent-sampling probe over a history sharded by particle: local (max, sum e, sum e^2) -> global triple,
    # identical on every rank and equal to the single-process reduction of the whole history
    logw = np.random.default_rng(5).normal(size=(6, n)) * 30.0            # [T, N] with a huge dynamic range
    mine_lw = logw[:, a:b].reshape(-1)
    mx = mine_lw.max()
    e = np.exp(mine_lw - mx)
    got = dist.combine_weight_stats(torch.tensor([mx, e.sum(), (e * e).sum()], dtype=torch.float64)).numpy()
    allw = logw.reshape(-1)
    eg = np.exp(allw - allw.max())
    want = np.array([allw.max(), eg.sum(), (eg * eg).sum()])
    ok &= got[0] == want[0] and bool(np.allclose(got[1:], want[1:], rtol=1e-13, atol=0.0))
    both = dist.gather_blocks(torch.from_numpy(got.reshape(1, 3).copy()))
    ok &= bool(torch.equal(both[0], both[1]))                              # bit-identical on the two ranks
    # scalars of the sharded history come back as [T, N] in global particle order; global trim / resampling
    # indices map to (owner, local flat index) and every selected row is found exactly once
    full = dist.gather_history_scalars(torch.from_numpy(logw[:, a:b].copy()), cnt)
    ok &= np.array_equal(full.numpy(), logw)
    pick = np.sort(np.random.default_rng(6).choice(logw.size, size=500, replace=False))
    owner, local = dist.split_history_index(pick, n, cnt)
    mine_sel = local[owner == rank]
    vals = torch.from_numpy(logw[:, a:b].reshape(-1)[mine_sel].copy()).reshape(-1, 1)
    sel_counts = [int((owner == r).sum()) for r in range(world)]
    back_vals = dist.gather_blocks(vals, sel_counts).numpy().reshape(-1)
    want_vals = np.concatenate([logw.reshape(-1)[pick[owner == r]] for r in range(world)])
    ok &= np.array_equal(back_vals, want_vals) and sum(sel_counts) == 500
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
    mgr = mp.get_context("spawn").Manager()      # never fork a multi-threaded (CUDA) parent
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_merge_weight_stats_rule():
    """the host mirror of csrc/smc_ops.cu Lse3::merge: order of the maxima, empty shards, associativity to rounding"""
    from pocomc_b200 import dist
    rng = np.random.default_rng(1)
    lw = rng.normal(size=5000) * 50.0
    cuts = [0, 17, 17, 2000, 4999, 5000]                                     # includes an empty shard
    parts = []
    for lo, hi in zip(cuts, cuts[1:]):
        seg = lw[lo:hi]
        if seg.size == 0:
            parts.append([-np.inf, 0.0, 0.0])
            continue
        e = np.exp(seg - seg.max())
        parts.append([seg.max(), e.sum(), (e * e).sum()])
    got = dist.merge_weight_stats(np.array(parts))
    e = np.exp(lw - lw.max())
    np.testing.assert_equal(got[0], lw.max())
    np.testing.assert_allclose(got[1:], [e.sum(), (e * e).sum()], rtol=1e-13)
    np.testing.assert_array_equal(dist.merge_weight_stats(np.array([[-np.inf, 0.0, 0.0]])), [-np.inf, 0.0, 0.0])
    # without a process group the combine is the identity
    t = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    assert dist.combine_weight_stats(t) is t


def test_split_history_index_arithmetic():
    from pocomc_b200 import dist
    n, T, counts = 10, 4, [3, 0, 5, 2]
    idx = np.arange(n * T)
    owner, local = dist.split_history_index(idx, n, counts)
    hist = np.arange(n * T).reshape(T, n)                                    # value = global flat index
    starts = np.concatenate([[0], np.cumsum(counts)])
    for r, c in enumerate(counts):
        shard = hist[:, starts[r]:starts[r] + c].reshape(-1)
        np.testing.assert_array_equal(shard[local[owner == r]], idx[owner == r])
    assert not np.any(owner == 1)                                            # the empty shard owns nothing
    import pytest
    with pytest.raises(ValueError):
        dist.split_history_index(idx, n, [3, 3])
