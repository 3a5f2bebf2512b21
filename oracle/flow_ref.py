"""TEST INFRASTRUCTURE -- CPU restatement of pocoMC's flow wrapper and trainer (pocomc/flow.py).

Flow arithmetic comes from the in-repo ``zuko`` restatement (oracle/zuko, parity unpinned against
real zuko -- see its header).  The wrapper / trainer logic here is pinned against the reference's
own pocomc/flow.py through tests/golden/flow*.npz (made by oracle/make_golden.py).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import zuko  # noqa: E402  (oracle/zuko)

PRESETS = {"maf3": ("maf", 3), "maf6": ("maf", 6), "maf12": ("maf", 12),
           "nsf3": ("nsf", 3), "nsf6": ("nsf", 6), "nsf12": ("nsf", 12)}


def hidden_width(n_dim: int) -> int:
    """flow.py:49-52: max(next_pow2(3 D), 32)."""
    n = 3 * n_dim
    p = 1 if n == 0 else 2 ** (n - 1).bit_length()
    return max(p, 32)


def make_flow(n_dim: int, preset: str = "nsf3"):
    """flow.py:54-86: zuko MAF/NSF with hidden_features=[H]*3, residual=True (NSF: bins=8)."""
    kind, t = PRESETS[preset]
    h = hidden_width(n_dim)
    if kind == "maf":
        return zuko.flows.MAF(n_dim, transforms=t, hidden_features=[h] * 3, residual=True)
    return zuko.flows.NSF(features=n_dim, bins=8, transforms=t, hidden_features=[h] * 3, residual=True)


def load_params(flow, arrays):
    """Copy a list of numpy arrays into the flow's parameters in module order."""
    with torch.no_grad():
        for p, a in zip(flow.parameters(), arrays):
            p.copy_(torch.as_tensor(np.asarray(a), dtype=p.dtype).reshape(p.shape))
    return flow


class NumpyFlow:
    """tools.flow_numpy_wrapper (tools.py:318-349): numpy f64 -> torch f32 -> flow -> numpy f32.
    forward NEGATES the ladj (tools.py:340); inverse does not (:348)."""

    def __init__(self, flow):
        self.flow = flow

    @torch.no_grad()
    def forward(self, v):
        t = torch.tensor(np.asarray(v), dtype=torch.float32)
        theta, ladj = self.flow().transform.call_and_ladj(t)
        return theta.numpy(), -ladj.numpy()

    @torch.no_grad()
    def inverse(self, theta):
        t = torch.tensor(np.asarray(theta), dtype=torch.float32)
        v, ladj = self.flow().transform.inv.call_and_ladj(t)
        return v.numpy(), ladj.numpy()


def fit(flow, x, weights=None, validation_split=0.0, epochs=1000, batch_size=1000, patience=20,
        learning_rate=1e-3, weight_decay=0, shuffle=True, clip_grad_norm=1.0, batches=None):
    """Flow.fit (flow.py:165-384) without the optional noise / annealing / regularisation branches
    (all off in Sampler's defaults, sampler.py:287-299).

    RNG consumption matches the reference: one global ``torch.randperm`` then a
    torch DataLoader per split per epoch.  ``batches`` (optional) replaces the DataLoaders with an
    explicit list, per epoch, of (train index batches, validation index batches) -- used by the
    parity tests to feed the CUDA trainer and this oracle identical mini-batches.
    """
    from torch.utils.data import DataLoader, TensorDataset
    x = x.float()
    n = x.shape[0]
    if shuffle:
        perm = torch.randperm(n)
        x = x[perm]
        if weights is not None:
            weights = weights[perm]
    n_train = int(validation_split * n) if validation_split > 0.0 else n
    validation = validation_split > 0.0
    opt = torch.optim.AdamW(flow.parameters(), learning_rate, weight_decay=weight_decay)

    def loss_of(idx_or_batch):
        if isinstance(idx_or_batch, (list, tuple)):
            xb = idx_or_batch[0]
            wb = idx_or_batch[1] if weights is not None else None
        else:
            xb = x[idx_or_batch]
            wb = weights[idx_or_batch] if weights is not None else None
        lp = flow().log_prob(xb)
        if wb is None:
            return -lp.sum()
        return (-lp * wb * 1000.0).sum() / wb.sum()

    if batches is None:
        sets = (x[:n_train],) if weights is None else (x[:n_train], weights[:n_train])
        train_dl = DataLoader(TensorDataset(*sets), batch_size, shuffle)
        if validation:
            vsets = (x[n_train:],) if weights is None else (x[n_train:], weights[n_train:])
            val_dl = DataLoader(TensorDataset(*vsets), batch_size, shuffle)
    history = dict(loss=[], val_loss=[])
    best_epoch, best_loss = 0, np.inf
    best = copy.deepcopy(flow.state_dict())
    for epoch in range(epochs):
        flow.train()
        tot = 0.0
        for b in (train_dl if batches is None else batches[epoch][0]):
            opt.zero_grad()
            loss = loss_of(b)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(flow.parameters(), clip_grad_norm)
            opt.step()
            tot += loss.data.item()
        history["loss"].append(tot / n_train)
        if validation:
            flow.eval()
            vt = 0.0
            for b in (val_dl if batches is None else batches[epoch][1]):
                vt += loss_of(b).data.item()
            history["val_loss"].append(vt / (n - n_train))
        mon = history["val_loss" if validation else "loss"][-1]
        if mon < best_loss:
            best_loss, best_epoch = mon, epoch
            best = copy.deepcopy(flow.state_dict())
        if epoch - best_epoch >= int(1.5 * patience):
            flow.load_state_dict(best)
            break
    return history
