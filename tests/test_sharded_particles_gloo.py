"""CPU, world_size 2 (gloo): pocomc_b200.sharded.ShardedParticles -- a history sharded by particle answers the
beta probe, the weights and the trimmed-row gather of Sampler._reweight exactly like the unsharded history
(SURVEY section 8e).  The device kernels are replaced by the numpy stand-ins of tests/fake_lib.py; what is under test
is the exchange: rank-ordered merge of the probe statistics, [T, N] reassembly of the weights, ownership of rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

T_ITERS, N, D = 7, 1000, 3


def _history(seed=4):
    rng = np.random.default_rng(seed)
    logl = rng.normal(size=(T_ITERS, N)) * 8.0 - 20.0
    u = rng.normal(size=(T_ITERS, N, D))
    beta = np.sort(np.concatenate([[0.0], rng.random(T_ITERS - 1)]))
    logz = np.cumsum(rng.normal(size=T_ITERS)) * 0.3
    logz[0] = 0.0
    return logl, u, beta, logz


def _fill(p, logl, u, beta, logz, cols=slice(None)):
    for t in range(T_ITERS):
        p.update(dict(logl=logl[t, cols].copy(), u=u[t, cols].copy(), x=2.0 * u[t, cols], logdetj=u[t, cols, 0] - 1.0,
                      logp=-0.5 * np.sum(u[t, cols] ** 2, axis=1), beta=float(beta[t]), logz=float(logz[t]), iter=t))


def _reweight_with(store):
    """pocomc_b200.sampler.Sampler._reweight on a bare namespace carrying only what the method reads"""
    import types
    from pocomc_b200.sampler import Sampler

    class _Bar:
        def update_stats(self, info): pass
        def update_iter(self): pass

    ns = types.SimpleNamespace(particles=store, t=T_ITERS, pbar=_Bar(), n_effective=400, n_active=N, dynamic=True,
                               dynamic_ratio=0.8, metric="ess", have_blobs=False)
    ns._probe = lambda beta: Sampler._probe(ns, beta)
    ns._ess_of_probe = lambda p: Sampler._ess_of_probe(ns, p)
    cur = Sampler._reweight(ns, {})
    cur["n_effective"] = ns.n_effective
    return cur


class _Patch:
    """minimal stand-in for pytest's monkeypatch inside spawned workers"""
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import fake_lib
    from pocomc_b200 import dist
    from pocomc_b200.particles import Particles
    from pocomc_b200.sharded import ShardedParticles
    fake_lib.install(_Patch())
    dist.init_from_env("gloo")
    logl, u, beta, logz = _history()
    counts = dist.shard_counts(N, world, 256)
    lo, hi = dist.shard_range(N, rank, world, 256)
    full = Particles(N, D)
    _fill(full, logl, u, beta, logz)
    mine = ShardedParticles(N, D, counts, rank)
    _fill(mine, logl, u, beta, logz, slice(lo, hi))
    ok = True
    for b in (0.0, 0.37, 1.0):
        pf, ps = full.probe(b), mine.probe(b)
        ok &= ps["m"] == pf["m"] == T_ITERS * N and ps["max"] == pf["max"]
        ok &= bool(np.isclose(ps["ess"], pf["ess"], rtol=1e-12)) and bool(np.isclose(ps["logz"], pf["logz"], rtol=1e-12, atol=1e-12))
        w_full = full.weights_device(b, stats=pf["stats"]).numpy()
        w_glob = mine.global_scalars(mine.weights_device(b, stats=ps["stats"])).numpy()
        ok &= w_glob.shape == w_full.shape and bool(np.allclose(w_glob, w_full, rtol=1e-12, atol=0.0))
        ok &= bool(np.isclose(w_glob.sum(), 1.0, rtol=1e-12))
    # the statistics are bit-identical on the two ranks (same branch of the bisection everywhere)
    both = dist.gather_blocks(mine.probe(0.5)["stats"].reshape(1, 4))
    ok &= bool(torch.equal(both[0], both[1]))
    # trimmed rows by global flat index, in global order, on every rank
    idx = np.sort(np.random.default_rng(8).choice(T_ITERS * N, size=900, replace=False))
    ok &= np.array_equal(mine.take_flat_global("u", idx), full.take_flat("u", idx))
    ok &= np.array_equal(mine.take_flat_global("logl", idx), full.take_flat("logl", idx))
    ok &= mine.take_flat_global("u", np.array([], dtype=np.int64)).shape == (0, D)
    # update() cuts whole-population arrays to this rank's block; reading the whole history back is a collective
    cut = ShardedParticles(N, D, counts, rank)
    _fill(cut, logl, u, beta, logz)                                   # N rows per iteration, like Sampler hands them over
    ok &= len(cut.past["logl"][0]) == hi - lo and np.array_equal(cut.past["u"][3], u[3, lo:hi])
    ok &= np.array_equal(cut.get("u"), full.get("u")) and np.array_equal(cut.get("u", flat=True), full.get("u", flat=True))
    ok &= np.array_equal(cut.get("logl", flat=True), full.get("logl", flat=True))
    ok &= cut.get("beta", index=-1) == full.get("beta", index=-1) and np.array_equal(cut.get("beta"), full.get("beta"))
    lw_s, z_s = cut.compute_logw_and_logz(1.0)
    lw_f, z_f = full.compute_logw_and_logz(1.0)
    ok &= lw_s.shape == lw_f.shape and bool(np.allclose(lw_s, lw_f, rtol=1e-12, atol=1e-12)) and bool(np.isclose(z_s, z_f, rtol=1e-12))
    res_s, res_f = cut.compute_results(), full.compute_results()
    ok &= set(res_s) == set(res_f) and np.array_equal(res_s["x"], res_f["x"]) and res_s["logw"].shape == res_f["logw"].shape
    # the whole of Sampler._reweight (bisection on beta, dynamic n_effective, trimming, gather of the survivors)
    # run on the sharded store must reproduce the run on the unsharded one, on every rank
    ra, rb = _reweight_with(full), _reweight_with(mine)
    ok &= ra["beta"] == rb["beta"] and ra["n_effective"] == rb["n_effective"]
    ok &= bool(np.isclose(ra["logz"], rb["logz"], rtol=1e-12, atol=1e-12)) and bool(np.isclose(ra["ess"], rb["ess"], rtol=1e-12))
    ok &= ra["weights"].shape == rb["weights"].shape and bool(np.allclose(ra["weights"], rb["weights"], rtol=1e-11, atol=0.0))
    for key in ("u", "x", "logdetj", "logl", "logp"):
        ok &= np.array_equal(ra[key], rb[key])
    ok &= 0.0 < ra["beta"] < 1.0 and 0 < len(ra["weights"]) < T_ITERS * N
    out[rank] = bool(ok)
    td.destroy_process_group()


def test_sharded_particles_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_sharded_particles_single_process_is_the_plain_history(monkeypatch):
    import fake_lib
    from pocomc_b200.particles import Particles
    from pocomc_b200.sharded import ShardedParticles
    fake_lib.install(monkeypatch)
    logl, u, beta, logz = _history(9)
    full, one = Particles(N, D), ShardedParticles(N, D, [N], 0)
    _fill(full, logl, u, beta, logz)
    _fill(one, logl, u, beta, logz)
    pf, ps = full.probe(0.6), one.probe(0.6)
    assert ps["ess"] == pytest.approx(pf["ess"], rel=1e-14) and ps["logz"] == pytest.approx(pf["logz"], rel=1e-14)
    # the fake denominators follow particles.py:222: logw = beta logl - (LSE_i(beta_i logl - logz_i) - log T)
    lw = logl * 0.6 - (np.logaddexp.reduce(beta[:, None, None] * logl[None] - logz[:, None, None], axis=0) - np.log(T_ITERS))
    assert pf["logz"] == pytest.approx(np.logaddexp.reduce(lw.reshape(-1)) - np.log(lw.size), rel=1e-12)
    with pytest.raises(NotImplementedError):
        one.probe(1.0, uss_k=10)
    with pytest.raises(ValueError):
        ShardedParticles(N, D, [N - 1], 0)


def _ckpt_worker(rank, world, port, tmp, out):
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import fake_lib
    from pocomc_b200 import dist
    from pocomc_b200.particles import Particles
    from pocomc_b200.sampler import Sampler
    from pocomc_b200.sharded import ShardedParticles
    fake_lib.install(_Patch())
    dist.init_from_env("gloo")
    logl, u, beta, logz = _history(11)
    counts = dist.shard_counts(N, world, 256)
    lo, hi = dist.shard_range(N, rank, world, 256)
    ok = True

    class _Bare:                                    # carries only what save_state / load_state touch
        pass

    def sampler_like(store):
        ns = _Bare()
        ns.__dict__.update(particles=store, t=T_ITERS, pbar=None, pool=None, distribute=None, calls=123)
        ns._state_path = types.MethodType(Sampler._state_path, ns)
        return ns

    # sharded history: every rank writes its own slice (<path>, <path>.rank1) and reads it back
    mine = ShardedParticles(N, D, counts, rank)
    _fill(mine, logl, u, beta, logz, slice(lo, hi))
    path = os.path.join(tmp, "sharded.state")
    Sampler.save_state(sampler_like(mine), path)
    ok &= os.path.exists(path) and os.path.exists(path + ".rank1") and not os.path.exists(path + ".temp")
    fresh = sampler_like(ShardedParticles(N, D, counts, rank))
    Sampler.load_state(fresh, path)
    ok &= fresh.t == T_ITERS and len(fresh.particles.past["logl"]) == T_ITERS
    ok &= all(np.array_equal(a, b) for a, b in zip(fresh.particles.past["u"], mine.past["u"]))
    ok &= np.array_equal(fresh.particles.get("logl", flat=True), logl.reshape(-1))          # the collective read sees the whole history
    ok &= bool(np.isclose(fresh.particles.probe(0.4)["ess"], mine.probe(0.4)["ess"], rtol=1e-12))
    # replicated history: rank 0 alone writes, everybody reads the same file
    full = Particles(N, D)
    _fill(full, logl, u, beta, logz)
    path2 = os.path.join(tmp, "replicated.state")
    Sampler.save_state(sampler_like(full), path2)
    ok &= os.path.exists(path2) and not os.path.exists(path2 + ".rank1")
    again = sampler_like(Particles(N, D))
    Sampler.load_state(again, path2)
    ok &= np.array_equal(again.particles.get("u"), full.get("u")) and again.calls == 123
    out[rank] = bool(ok)
    td.barrier()
    td.destroy_process_group()


def test_checkpoint_of_sharded_and_replicated_history_world2(tmp_path):
    """SURVEY 8(f4), sampler.py:1023-1061 under torch.distributed: a sharded history is saved as one file per rank and
    resumes on the same sharding; a replicated state is written once, by rank 0."""
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_ckpt_worker, args=(world, _free_port(), str(tmp_path), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
