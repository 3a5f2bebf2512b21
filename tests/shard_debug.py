"""Diagnostic (not a test): per-iteration checksums of a single-process and a 2-rank sharded run."""
import os, sys, socket, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PMC_SWEEP_LPP"] = "4"

def loglike(x):
    return -0.5 * np.sum(((x - 0.5) / 0.3) ** 2, axis=1) - 5.0 * (x[:, 0] ** 2 - x[:, 1]) ** 2

def run(tag):
    from scipy.stats import norm, uniform
    import pocomc_b200 as pc
    from pocomc_b200 import config
    config.set_rng_mode("host"); config.mean_mode = 0
    log = []
    S = pc.Sampler
    orig_train, orig_mut, orig_rew, orig_res = S._train, S._mutate, S._reweight, S._resample
    def tr(self, cp):
        out = orig_train(self, cp)
        blob = self.flow.flow.raw.detach().double().sum().item()
        log.append(("train", self.t, blob, float(self.theta_geometry.t_mean.sum()) if self.theta_geometry.t_mean is not None else 0.0))
        return out
    def mu(self, cp):
        out = orig_mut(self, cp)
        log.append(("mutate", self.t, float(out["x"].sum()), float(out["logl"].sum()), int(out["steps"]), float(out["accept"]), float(self.proposal_scale)))
        return out
    def rw(self, cp):
        out = orig_rew(self, cp)
        log.append(("reweight", self.t, float(out["beta"]), float(out["x"].sum()), float(out["weights"].sum()), len(out["weights"])))
        return out
    def rs(self, cp):
        out = orig_res(self, cp)
        log.append(("resample", self.t, float(out["x"].sum())))
        return out
    S._train, S._mutate, S._reweight, S._resample = tr, mu, rw, rs
    prior = pc.Prior([uniform(-3.0, 6.0), norm(0.0, 2.0), uniform(-3.0, 6.0), norm(0.0, 2.0)])
    s = pc.Sampler(prior, loglike, vectorize=True, n_active=512, n_effective=1024, flow="maf3",
                   train_config=dict(epochs=30), random_state=3)
    s.run(n_total=1024, n_evidence=512, progress=False)
    S._train, S._mutate, S._reweight, S._resample = orig_train, orig_mut, orig_rew, orig_res
    return log

def worker(rank, world, port, out):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    from pocomc_b200 import dist
    dist.init_from_env("gloo")
    out[rank] = run(f"r{rank}")
    torch.distributed.barrier()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    a = run("single1")
    b = run("single2")
    mgr = mp.Manager(); out = mgr.dict()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    c, d = out[0], out[1]
    for i, (x, y, z, w) in enumerate(zip(a, b, c, d)):
        flag = ("" if x == y else " RUN2RUN") + ("" if x == z else " SHARD0") + ("" if z == w else " RANKS")
        print(i, x, flag)
        if flag:
            print("   single2:", y); print("   rank0  :", z); print("   rank1  :", w)
            break
