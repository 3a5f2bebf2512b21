"""GPU, world_size 2: a particle-sharded ``Sampler.run()`` (one process per rank)
must reproduce the single-process run -- same temperature ladder, same logZ, same final particles --
because every rank walks the same host RNG stream, the per-step reductions are fixed-order block
partials and the per-temperature all-gather returns the mutated rows in global particle order.
On a box with fewer than two GPUs both ranks share cuda:0 and talk over gloo (host-staged
collectives); with two or more GPUs the ranks use NCCL.  Bar: bit-exact (the flow sweep's K-slicing
is pinned with PMC_SWEEP_LPP so the fp32 summation order does not depend on the shard size)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_ACTIVE, N_DIM = 512, 4


def _loglike(x):
    return -0.5 * np.sum(((x - 0.5) / 0.3) ** 2, axis=1) - 5.0 * (x[:, 0] ** 2 - x[:, 1]) ** 2


def _run(sample="tpcn", precondition=True):
    from scipy.stats import norm, uniform
    import pocomc_b200 as pc
    from pocomc_b200 import config
    config.set_rng_mode("host")
    config.mean_mode = 0                       # the GPU-count independent block-partial mean (sharded runs always use it)
    prior = pc.Prior([uniform(-3.0, 6.0), norm(0.0, 2.0), uniform(-3.0, 6.0), norm(0.0, 2.0)])
    s = pc.Sampler(prior, _loglike, vectorize=True, n_active=N_ACTIVE, n_effective=2 * N_ACTIVE, flow="maf3",
                   train_config=dict(epochs=30), random_state=3, sample=sample, precondition=precondition)
    s.run(n_total=1024, n_evidence=512, progress=False)
    r = s.results
    return dict(logz=float(s.evidence()[0]), beta=np.asarray(r["beta"]), x=np.asarray(r["x"][-1]),
                logl=np.asarray(r["logl"][-1]), calls=int(s.calls), steps=np.asarray(r["steps"]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, kwargs, out):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      PMC_SWEEP_LPP="4")
    ngpu = torch.cuda.device_count()
    os.environ["LOCAL_RANK"] = str(rank % ngpu)
    torch.cuda.set_device(rank % ngpu)
    from pocomc_b200 import dist
    dist.init_from_env("nccl" if ngpu >= world else "gloo")
    res = _run(**kwargs)
    out[rank] = res
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("kwargs", [dict(sample="tpcn", precondition=True), dict(sample="rwm", precondition=False)])
def test_sharded_run_equals_single_process(kwargs):
    import torch.multiprocessing as mp
    os.environ["PMC_SWEEP_LPP"] = "4"
    try:
        single = _run(**kwargs)
    finally:
        from pocomc_b200 import config
        config.mean_mode = None
        os.environ.pop("PMC_SWEEP_LPP", None)
    world = 2
    mgr = mp.get_context("spawn").Manager()      # never fork a multi-threaded (CUDA) parent
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kwargs, out), nprocs=world, join=True)
    for r in range(world):
        got = out[r]
        np.testing.assert_array_equal(got["beta"], single["beta"])
        np.testing.assert_array_equal(got["steps"], single["steps"])
        np.testing.assert_array_equal(got["x"], single["x"])
        np.testing.assert_array_equal(got["logl"], single["logl"])
        assert got["calls"] == single["calls"]
        assert got["logz"] == single["logz"]


def test_sharded_history_run_equals_single_process(monkeypatch):
    """Same bar as above with every rank storing only its block of the history (probe statistics merged in rank
    order, weights gathered as scalars, trimmed rows gathered by owner)."""
    import torch.multiprocessing as mp
    kwargs = dict(sample="tpcn", precondition=True)
    os.environ["PMC_SWEEP_LPP"] = "4"
    try:
        single = _run(**kwargs)
    finally:
        from pocomc_b200 import config
        config.mean_mode = None
        os.environ.pop("PMC_SWEEP_LPP", None)
    monkeypatch.setenv("PMC_B200_SHARD_HISTORY", "1")             # read by pocomc_b200.config in the spawned workers
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kwargs, out), nprocs=world, join=True)
    for r in range(world):
        got = out[r]
        np.testing.assert_array_equal(got["beta"], single["beta"])
        np.testing.assert_array_equal(got["steps"], single["steps"])
        np.testing.assert_array_equal(got["x"], single["x"])
        assert got["calls"] == single["calls"]
        np.testing.assert_allclose(got["logz"], single["logz"], rtol=1e-12)

