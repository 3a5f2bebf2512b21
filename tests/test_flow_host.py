"""CPU: host-side logic of pocomc_b200.flow that needs no device."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

import flow_ref as F
from pocomc_b200.flow import Flow, epoch_batches


@pytest.mark.parametrize("preset,d", [("maf3", 4), ("nsf6", 7), ("maf12", 2)])
def test_init_matches_zuko_rng_order(preset, d):
    """Same torch seed -> same initial weights as constructing the zuko flow."""
    torch.manual_seed(123)
    ref = F.make_flow(d, preset)
    torch.manual_seed(123)
    mine = Flow(d, preset)
    flat = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    assert torch.equal(flat, mine.flow.raw.detach().cpu())
    # and the generator is left in the same state
    assert torch.equal(torch.rand(3), (torch.manual_seed(123), F.make_flow(d, preset), torch.rand(3))[2])


def test_invalid_flow_name():
    with pytest.raises(ValueError):
        Flow(4, "realnvp")
    with pytest.raises(ValueError):
        Flow(1, "maf3")           # 1-D autoregressive net has a null Jacobian (zuko raises too)


@pytest.mark.parametrize("shuffle", [True, False])
def test_epoch_batches_replicates_dataloader(shuffle):
    n, bs = 103, 16
    data = torch.arange(n)
    torch.manual_seed(5)
    dl = DataLoader(TensorDataset(data), bs, shuffle)
    ref = [[b[0] for b in dl] for _ in range(3)]
    after_ref = torch.rand(2)
    torch.manual_seed(5)
    mine = [epoch_batches(n, bs, shuffle) for _ in range(3)]
    after = torch.rand(2)
    for e_ref, e in zip(ref, mine):
        assert len(e_ref) == len(e)
        for a, b in zip(e_ref, e):
            assert torch.equal(a, b)
    assert torch.equal(after_ref, after)


def test_plateau_lr_mirrors_torch_scheduler():
    """_PlateauLR (used by the CUDA-graph fit path) follows torch's ReduceLROnPlateau(mode='min',
    factor=0.2, threshold=1e-4, threshold_mode='abs', min_lr=1e-6) decision for decision."""
    import numpy as np
    import torch
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    from pocomc_b200.flow import _PlateauLR
    rng = np.random.default_rng(5)
    losses = np.concatenate([np.linspace(3, 1, 15), 1 + 0.01 * rng.normal(size=60), np.linspace(1, 0.5, 10), np.full(40, 0.5)])
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], 1e-3)
    ref = ReduceLROnPlateau(opt, mode='min', factor=0.2, patience=3, threshold=0.0001, threshold_mode='abs', min_lr=1e-6)
    mine = _PlateauLR(1e-3, factor=0.2, patience=3, threshold=0.0001, min_lr=1e-6)
    for v in losses:
        ref.step(float(v))
        mine.step(float(v))
        assert abs(opt.param_groups[0]['lr'] - mine.lr) < 1e-15
    assert mine.lr < 1e-3


@pytest.mark.parametrize("kind,d,h", [("maf", 5, 48), ("nsf", 6, 40)])
def test_adopts_a_ready_zuko_flow(kind, d, h):
    """pocomc/flow.py:87-88: ``Flow(n_dim, flow=<zuko.flows.Flow>)``.  The module's parameters land in the flat blob in
    module order, the structure (width, transforms, head) is read off the module, and the trained blob can be written back."""
    import zuko
    torch.manual_seed(4)
    kw = dict(transforms=4, hidden_features=[h] * 3, residual=True)
    user = zuko.flows.MAF(d, **kw) if kind == "maf" else zuko.flows.NSF(features=d, bins=8, **kw)
    mine = Flow(d, user)
    lay = mine.flow.layout
    assert (lay.n_dim, lay.n_hidden, lay.n_layers, lay.n_transforms) == (d, h, 3, 4)
    flat = torch.cat([p.detach().reshape(-1) for p in user.parameters()])
    assert torch.equal(flat, mine.flow.raw.detach().cpu())
    with torch.no_grad():
        mine.flow.raw.add_(1.0)
    mine.flow.export_to(user)
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in user.parameters()]), flat + 1.0)


def test_rejects_zuko_flows_the_kernels_do_not_implement():
    import zuko
    with pytest.raises(ValueError, match="residual"):
        Flow(4, zuko.flows.MAF(4, transforms=2, hidden_features=[32] * 3, residual=False))
    with pytest.raises(ValueError, match="3 hidden layers"):
        Flow(4, zuko.flows.MAF(4, transforms=2, hidden_features=[32] * 2, residual=True))
    with pytest.raises(ValueError, match="orders must alternate"):
        Flow(4, zuko.flows.MAF(4, transforms=2, hidden_features=[32] * 3, residual=True, randperm=True))
    with pytest.raises(ValueError, match="bins = 8"):
        Flow(4, zuko.flows.NSF(features=4, bins=4, transforms=2, hidden_features=[32] * 3, residual=True))
    with pytest.raises(ValueError):
        Flow(4, torch.nn.Linear(4, 4))
