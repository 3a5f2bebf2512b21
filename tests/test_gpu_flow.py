"""GPU parity: the sm_100a flow sweep kernels (through the C-ABI) against the zuko oracle and the
golden vectors recorded from the reference's Flow wrapper.  fp32 tolerance: 2e-5 abs/rel for
maf, 2e-4 for nsf (spline knots amplify rounding), stated per assert."""
import numpy as np
import pytest
import torch

import flow_ref as F
from conftest import flow_param_list

pytestmark = pytest.mark.gpu


def _bar(ref, base=5e-5):
    """fp32 parity bar of a flow output: relative, with the absolute part scaled by the magnitude of the reference values
    (a chain of 24-48 fp32 layers with exp() scales carries rounding noise proportional to the values it transports)."""
    return dict(rtol=base, atol=base * max(1.0, float(torch.as_tensor(ref).abs().max())))


def _mine(preset, d, params):
    from pocomc_b200.flow import Flow
    f = Flow(d, preset)
    flat = np.concatenate([np.asarray(p).reshape(-1) for p in params])
    with torch.no_grad():
        f.flow.raw.copy_(torch.from_numpy(flat).to(f.flow.raw.device))
    return f


@pytest.mark.parametrize("preset,d", [("maf3", 4), ("nsf3", 5), ("maf6", 10)])
def test_flow_matches_reference_wrapper_goldens(golden, preset, d):
    g = golden("flow")
    pre = preset + "_"
    f = _mine(preset, d, flow_param_list(g, pre))
    tol = dict(rtol=2e-5, atol=2e-5) if preset.startswith("maf") else dict(rtol=2e-4, atol=2e-4)
    x = torch.tensor(g[pre + "x"])
    with torch.no_grad():
        z, ladj = f.forward(x)
        xi, li = f.inverse(x)
        lp = f.log_prob(x)
    assert z.device == x.device and z.dtype == torch.float32
    np.testing.assert_allclose(z.numpy(), g[pre + "z"], **tol)
    np.testing.assert_allclose(ladj.numpy(), g[pre + "ladj"], **tol)
    np.testing.assert_allclose(xi.numpy(), g[pre + "inv_x"], **tol)
    np.testing.assert_allclose(li.numpy(), g[pre + "inv_ladj"], **tol)
    np.testing.assert_allclose(lp.numpy(), g[pre + "logprob"], **tol)
    # Flow.sample consumes the same host normals as zuko's base.rsample
    torch.manual_seed(8)
    with torch.no_grad():
        xs, lq = f.sample(32)
    np.testing.assert_allclose(xs.numpy(), g[pre + "sample_x"], **tol)
    np.testing.assert_allclose(lq.numpy(), g[pre + "sample_logq"], **tol)


@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("nsf6", 10, 257), ("maf3", 2, 33), ("nsf3", 3, 1),
                                        ("maf12", 50, 300), ("maf6", 100, 64), ("nsf6", 32, 500)])
def test_sweep_vs_oracle_sizes(preset, d, n):
    """every lanes-per-particle variant / ragged tile / big-H case against the oracle's D+1-pass inverse"""
    torch.manual_seed(d * 7 + n)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, l_ref = ref().transform.call_and_ladj(x)
        xi_ref, li_ref = ref().transform.inv.call_and_ladj(z_ref)
        z, l = f.forward(x)
        xi, li = f.inverse(z_ref)
    base = 5e-5 if preset.startswith("maf") else 5e-4
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **_bar(z_ref, base))
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), **_bar(l_ref, base))
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), **_bar(xi_ref, base))
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), **_bar(li_ref, base))


@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("maf12", 50, 300)])
def test_ffma_stream_variant_matches_oracle(preset, d, n):
    """the default fp32-FMA TMA-stream sweep, both directions through the sweep kernel (Flow.forward itself takes tcgen05)"""
    torch.manual_seed(d + 2 * n)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    assert int(f.flow._meta_host[22]) == 2
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, l_ref = ref().transform.call_and_ladj(x)
        xi_ref, li_ref = ref().transform.inv.call_and_ladj(z_ref)
        z, l = f.flow.sweep(x, inverse=False)
        xi, li = f.flow.sweep(z_ref, inverse=True)
    tol = dict(rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), **tol)
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), **tol)
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), **tol)


def test_flow_properties_like_reference_tests():
    """reference tests/test_flow.py: round trip 1e-5, ladj antisymmetry, f64 warning, 1-row batch, fit stays finite"""
    from pocomc_b200.flow import Flow
    torch.manual_seed(0)
    x = torch.randn(100, 4) * 1.5
    f = Flow(4, "maf3")
    with torch.no_grad():
        z, ladj = f.forward(x)
        xr, li = f.inverse(z)
    assert torch.allclose(x, xr, atol=1e-5)
    torch.testing.assert_close(ladj, -li, atol=1e-5, rtol=1e-5)
    with pytest.warns(UserWarning):
        lp = f.log_prob(x.double())
    assert lp.dtype == torch.float32 and lp.shape == (100,) and torch.isfinite(lp).all()
    with pytest.raises(ValueError):
        f.forward(x.to(torch.int64))
    with torch.no_grad():
        z1, l1 = f.forward(x[:1])
    assert z1.shape == (1, 4) and l1.shape == (1,)
    # every parameter receives a gradient through log_prob (test_flow.py:136-150)
    f.flow.zero_grad()
    f.log_prob(x).sum().backward()
    g = f.flow.raw.grad
    assert g is not None and torch.isfinite(g).all()
    hist = f.fit(x, epochs=5)
    assert len(hist["loss"]) == 5
    with torch.no_grad():
        z, _ = f.forward(x)
        s, lq = f.sample(10)
    assert torch.isfinite(z).all() and torch.isfinite(s).all() and torch.isfinite(lq).all()


@pytest.mark.parametrize("preset,d", [("maf3", 4), ("nsf3", 6)])
def test_autograd_path_matches_oracle_gradients(preset, d):
    torch.manual_seed(3)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    x = torch.randn(64, d)
    lp_ref = ref().log_prob(x)
    lp_ref.sum().backward()
    gref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    lp = f.log_prob(x)
    lp.sum().backward()
    np.testing.assert_allclose(lp.detach().cpu().numpy(), lp_ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(f.flow.raw.grad.cpu().numpy(), gref.numpy(), rtol=1e-3, atol=1e-3)
    # masked weight entries never receive gradient
    with torch.no_grad():
        z_sweep, l_sweep = f.forward(x)
    z_auto, l_auto = f.flow.forward_autograd(x)
    np.testing.assert_allclose(z_sweep.numpy(), z_auto.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("tag", ["w", "nw"])
def test_fit_matches_reference_history(golden, tag):
    """Flow.fit against the loss history / final weights the reference produced from the same
    initial weights, data and torch seed (flow.py:165-384)."""
    from pocomc_b200.flow import Flow
    g = golden("flow_fit")
    f = _mine("maf3", 4, flow_param_list(g, f"{tag}_init_"))
    w = torch.tensor(g["w"]) if tag == "w" else None
    torch.manual_seed(14)
    hist = f.fit(torch.tensor(g["data"]), weights=w, validation_split=0.5, epochs=4, batch_size=64, patience=4,
                 annealing=False, shuffle=True, clip_grad_norm=1.0)
    np.testing.assert_allclose(hist["loss"], g[f"{tag}_loss"], rtol=2e-4)
    np.testing.assert_allclose(hist["val_loss"], g[f"{tag}_val_loss"], rtol=2e-4)
    final = np.concatenate([p.reshape(-1) for p in flow_param_list(g, f"{tag}_final_")])
    np.testing.assert_allclose(f.flow.raw.detach().cpu().numpy(), final, rtol=2e-2, atol=2e-4)


@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("maf3", 4, 33), ("maf6", 10, 257), ("maf12", 21, 128),
                                        ("maf3", 42, 129), ("maf6", 32, 20000), ("maf3", 2, 1)])
def test_tensor_core_forward_matches_oracle(preset, d, n):
    """csrc/flow_tc.cu (tcgen05, A operand in TMEM, 3xTF32 split) against the oracle's dense forward and
    against the FFMA sweep kernel: fp32 parity bar 5e-5 like the sweep (tolerance stated here)."""
    from pocomc_b200 import config
    torch.manual_seed(d * 11 + n)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    assert f.flow.tc_available()
    x = torch.randn(n, d) * 1.3
    with torch.no_grad():
        z_ref, l_ref = ref().transform.call_and_ladj(x)
    xd = x.cuda()
    z = torch.empty_like(xd)
    l = torch.empty(n, dtype=torch.float32, device="cuda")
    f.flow.forward_tc_into(xd, z, l, passes=3)
    tol = _bar(z_ref)
    np.testing.assert_allclose(z.cpu().numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(l.cpu().numpy(), l_ref.numpy(), **_bar(l_ref))
    zs, ls = f.flow.sweep(xd, False) if config.forward_path == "sweep" else (None, None)
    old = config.forward_path
    config.forward_path = "sweep"
    try:
        zs, ls = f.flow.sweep(xd, False)
    finally:
        config.forward_path = old
    np.testing.assert_allclose(z.cpu().numpy(), zs.cpu().numpy(), **tol)
    np.testing.assert_allclose(l.cpu().numpy(), ls.cpu().numpy(), **tol)
    # plain TF32 (passes = 1) is only TF32-accurate: loose bar, documents the precision trade
    f.flow.forward_tc_into(xd, z, l, passes=1)
    assert np.max(np.abs(z.cpu().numpy() - z_ref.numpy())) < 0.05 * max(1.0, float(z_ref.abs().max()))


@pytest.mark.parametrize("weighted,annealing", [(True, False), (False, True)])
def test_graph_fit_matches_eager_fit(weighted, annealing):
    """The CUDA-graph optimiser step (fused clip + AdamW kernel, padded ragged batches) reproduces the
    eager torch.optim.AdamW / clip_grad_norm_ path: same loss history to 1e-5, same weights to 1e-5."""
    from pocomc_b200 import config
    from pocomc_b200.flow import Flow
    torch.manual_seed(2)
    data = torch.randn(700, 5) * torch.tensor([1.0, 2.0, 0.5, 1.5, 1.0]) + 0.3
    w = torch.rand(700) + 0.1 if weighted else None
    hist, final = {}, {}
    for path in ("eager", "graph"):
        config.fit_path = path
        try:
            torch.manual_seed(7)
            f = Flow(5, "maf3")
            torch.manual_seed(9)
            hist[path] = f.fit(data, weights=w, validation_split=0.6, epochs=12, batch_size=128, patience=1 if annealing else 30,
                               annealing=annealing, shuffle=True, clip_grad_norm=1.0, learning_rate=2e-3)
            final[path] = f.flow.raw.detach().cpu().numpy().copy()
        finally:
            config.fit_path = "graph"
    np.testing.assert_allclose(hist["graph"]["loss"], hist["eager"]["loss"], rtol=1e-5)
    np.testing.assert_allclose(hist["graph"]["val_loss"], hist["eager"]["val_loss"], rtol=1e-5)
    np.testing.assert_allclose(final["graph"], final["eager"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("preset,d,n,weighted", [("maf3", 4, 100, True), ("maf6", 10, 512, False), ("maf6", 32, 300, True),
                                                 ("maf3", 42, 64, True), ("maf3", 21, 33, False),
                                                 ("maf6", 50, 512, True), ("maf3", 64, 96, False)])     # H = 256: images stream in k-chunks
def test_fused_training_kernels_match_oracle_gradients(preset, d, n, weighted):
    """csrc/flow_train.cu (fused forward + input-gradient chain, grouped weight-gradient GEMM) against the
    oracle's autograd: loss to 1e-5 relative, every parameter gradient to 1e-4 of the gradient scale."""
    from pocomc_b200.flow import _FitEngine
    torch.manual_seed(d + n)
    ref = F.make_flow(d, preset)
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    x = torch.randn(n + 50, d) * 1.2 + 0.1
    w = torch.rand(n + 50) + 0.05
    rows = torch.randperm(n + 50)[:n]
    lp = ref().log_prob(x[rows])
    if weighted:
        loss_ref = (-lp * w[rows] * 1000.0).sum() / w[rows].sum()
    else:
        loss_ref = -lp.sum()
    loss_ref.backward()
    gref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).numpy()
    eng = _FitEngine(f.flow)
    assert eng.fused
    eng.load(x.cuda(), w.cuda())
    loss, g = eng.loss_and_grad(rows, weighted)
    np.testing.assert_allclose(loss, float(loss_ref), rtol=1e-5)
    g = g.cpu().numpy()
    scale = np.abs(gref).max()
    np.testing.assert_allclose(g, gref, rtol=2e-4, atol=1e-4 * scale)
    # masked entries of the blob receive exactly zero gradient
    assert np.all(g[gref == 0] == 0)


@pytest.mark.parametrize("preset,d,n_fwd,n_inv", [("maf6", 100, 1000, 200), ("maf3", 200, 1000, 64), ("maf6", 50, 2000, 500)])
def test_config_shapes_vs_oracle(preset, d, n_fwd, n_inv):
    """BASELINE configs[2..4] flow shapes (D, H) = (50, 256), (100, 512), (200, 1024) against the oracle itself (not only
    the round-trip property): forward at n >= 1000 rows, and the oracle's true D+1-pass inverse on a smaller batch
    (606 / 603 dense hyper-network passes on the host).  Weights scaled away from the near-identity initialisation."""
    torch.manual_seed(d + 11)
    ref = F.make_flow(d, preset)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.mul_(1.2)
    f = _mine(preset, d, [p_.detach().numpy() for p_ in ref.parameters()])
    x = torch.randn(n_fwd, d)
    with torch.no_grad():
        z_ref, l_ref = ref().transform.call_and_ladj(x)
        z, l = f.forward(x)
        zi = z_ref[:n_inv].contiguous()
        xi_ref, li_ref = ref().transform.inv.call_and_ladj(zi)
        xi, li = f.inverse(zi)
    tol = dict(rtol=5e-5, atol=5e-5 * max(1.0, float(z_ref.abs().max())))
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), rtol=5e-5, atol=5e-5 * max(1.0, float(l_ref.abs().max())))
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), rtol=5e-5, atol=5e-5 * max(1.0, float(xi_ref.abs().max())))
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), rtol=5e-5, atol=5e-5 * max(1.0, float(li_ref.abs().max())))


@pytest.mark.parametrize("preset,d,n", [("maf6", 32, 1000), ("maf3", 10, 77), ("maf3", 21, 300), ("maf3", 16, 129),
                                        ("maf6", 33, 515), ("maf6", 32, 20011), ("maf12", 14, 64), ("maf3", 9, 200), ("maf3", 17, 130), ("maf3", 36, 257), ("maf3", 8, 100),
                                        ("maf3", 40, 300), ("maf6", 50, 1500), ("maf3", 100, 19000), ("maf3", 200, 300),
                                        ("nsf6", 10, 300), ("nsf6", 8, 77), ("nsf3", 32, 1000), ("nsf3", 21, 300), ("nsf3", 14, 130),
                                        ("nsf6", 32, 5011), ("nsf3", 50, 700), ("nsf3", 100, 400), ("nsf3", 200, 200)])
def test_tensor_core_block_triangular_sweep_matches_oracle(preset, d, n):
    """csrc/flow_tri.cu (tcgen05 right-looking block updates + in-block fp32 substitution), BOTH directions, against the
    oracle's 1-pass forward / D+1-pass inverse: ragged last tile, several tiles per CTA, every block shape (4, 5, 6 units
    per degree group), one tensor-memory window (D <= 36) and several (the BASELINE widths 50 / 100 / 200 with their scratch
    area); fp32 bar 5e-5.  It is the default path of Flow.inverse for these shapes."""
    from pocomc_b200 import config, made_layout as ML, tri_layout as TL
    spline = preset.startswith("nsf")          # zuko NSF (the reference's default presets): 24 output columns per order position + the RQS head
    assert TL.tri_supported(d, F.hidden_width(d), 3, ML.KIND_RQS if spline else ML.KIND_AFFINE)
    torch.manual_seed(d * 5 + n)
    ref = F.make_flow(d, preset)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.mul_(1.0 if preset == "maf12" else 1.25)        # 12 SCALED transforms are a chaotic map: rounding noise explodes
    f = _mine(preset, d, [p_.detach().numpy() for p_ in ref.parameters()])
    assert f.flow.tri_available() and config.inverse_path == "tri"
    x = torch.randn(n, d)
    m_ = min(n, 2000 if d <= 50 else 300)                       # the oracle's inverse is D+1 passes: bound its batch
    with torch.no_grad():
        z_ref, l_ref = ref().transform.call_and_ladj(x)
        xi_ref, li_ref = ref().transform.inv.call_and_ladj(z_ref[:m_])
    dev = f.flow.raw.device
    out = torch.empty(n, d, device=dev); ladj = torch.empty(n, device=dev)
    f.flow.sweep_tri_into(x.to(dev), out, ladj, inverse=False)
    # 12 scaled transforms amplify fp32 rounding with the magnitude of the values: the absolute bar scales with it
    bar = 5e-4 if spline else 5e-5                               # the spline flows' bar
    tol = dict(rtol=bar, atol=bar * max(1.0, float(z_ref.abs().max()), float(xi_ref.abs().max())))
    np.testing.assert_allclose(out.cpu().numpy(), z_ref.numpy(), **tol)
    np.testing.assert_allclose(ladj.cpu().numpy(), l_ref.numpy(), **tol)
    xi, li = f.inverse(z_ref)                                    # the public call takes the tensor-core path
    np.testing.assert_allclose(xi.numpy()[:m_], xi_ref.numpy(), **tol)
    np.testing.assert_allclose(li.numpy()[:m_], li_ref.numpy(), **tol)
    # in-place call (in == out) and agreement with the fp32-FMA sweep on every row
    buf = z_ref.to(dev).clone()
    f.flow.sweep_tri_into(buf, buf, ladj, inverse=True)
    np.testing.assert_array_equal(buf.cpu().numpy(), xi.numpy())
    old = config.inverse_path
    config.inverse_path = "sweep"
    try:
        xs, ls = f.inverse(z_ref)
    finally:
        config.inverse_path = old
    np.testing.assert_allclose(xi.numpy(), xs.numpy(), **tol)
    np.testing.assert_allclose(li.numpy(), ls.numpy(), **tol)


@pytest.mark.parametrize("kind,d,h", [("maf", 12, 64), ("nsf", 6, 40)])
def test_adopted_zuko_flow_matches_its_own_arithmetic(kind, d, h):
    """pocomc/flow.py:87-88, the inner plugin seam: a user-built zuko flow handed to Flow runs on the kernels and returns
    what the module itself computes (forward, inverse, log_prob); fit() writes the trained parameters back into it."""
    import zuko
    from pocomc_b200.flow import Flow
    torch.manual_seed(21)
    kw = dict(transforms=3, hidden_features=[h] * 3, residual=True)
    user = zuko.flows.MAF(d, **kw) if kind == "maf" else zuko.flows.NSF(features=d, bins=8, **kw)
    f = Flow(d, user)
    x = torch.randn(300, d)
    with torch.no_grad():
        z_ref, l_ref = user().transform.call_and_ladj(x)
        xi_ref, li_ref = user().transform.inv.call_and_ladj(z_ref)
        lp_ref = user().log_prob(x)
        z, l = f.forward(x)
        xi, li = f.inverse(z_ref)
        lp = f.log_prob(x)
    base = 5e-5 if kind == "maf" else 5e-4
    np.testing.assert_allclose(z.numpy(), z_ref.numpy(), **_bar(z_ref, base))
    np.testing.assert_allclose(l.numpy(), l_ref.numpy(), **_bar(l_ref, base))
    np.testing.assert_allclose(xi.numpy(), xi_ref.numpy(), **_bar(xi_ref, base))
    np.testing.assert_allclose(li.numpy(), li_ref.numpy(), **_bar(li_ref, base))
    np.testing.assert_allclose(lp.numpy(), lp_ref.numpy(), **_bar(lp_ref, base))
    before = torch.cat([p.detach().reshape(-1) for p in user.parameters()]).clone()
    f.fit(x, epochs=3, batch_size=100)
    after = torch.cat([p.detach().reshape(-1) for p in user.parameters()])
    assert torch.equal(after, f.flow.raw.detach().cpu()) and not torch.equal(after, before)
    with torch.no_grad():
        np.testing.assert_allclose(f.log_prob(x).numpy(), user().log_prob(x).numpy(), **_bar(lp_ref, 10 * base))


@pytest.mark.parametrize("preset,d,n,weighted,scale", [("nsf3", 5, 100, True, 1.0), ("nsf6", 10, 512, False, 1.0), ("nsf6", 10, 300, True, 3.0),
                                                       ("nsf3", 32, 77, True, 1.5), ("maf3", 100, 64, True, 1.0), ("maf3", 70, 200, False, 1.2),
                                                       ("maf3", 200, 33, True, 1.0), ("nsf3", 50, 40, False, 1.0)])
def test_layerwise_training_kernels_match_oracle_gradients(preset, d, n, weighted, scale):
    """csrc/flow_train_lw.cu (masked-linear GEMMs + affine / spline head kernels, forward and backward) against the oracle's
    autograd for the flows the fused kernels do not cover: spline flows (the reference's default presets), H = 512 / 1024,
    D > 64.  Data scaled so that spline inputs fall inside AND outside the [-5, 5] spline box.  Loss 1e-5 relative, every
    parameter gradient to 2e-4 of the gradient scale, masked entries exactly zero."""
    from pocomc_b200.flow import _FitEngine
    torch.manual_seed(d + n)
    ref = F.make_flow(d, preset)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.mul_(1.3)                       # away from the near-identity initialisation: knots and slopes that differ per feature
    f = _mine(preset, d, [p.detach().numpy() for p in ref.parameters()])
    x = torch.randn(n + 50, d) * 1.6 * scale + 0.1
    w = torch.rand(n + 50) + 0.05
    rows = torch.randperm(n + 50)[:n]
    lp = ref().log_prob(x[rows])
    loss_ref = (-lp * w[rows] * 1000.0).sum() / w[rows].sum() if weighted else -lp.sum()
    loss_ref.backward()
    gref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).numpy()
    eng = _FitEngine(f.flow)
    assert eng.layerwise and not eng.fused
    eng.load(x.cuda(), w.cuda())
    loss, g = eng.loss_and_grad(rows, weighted)
    np.testing.assert_allclose(loss, float(loss_ref.detach()), rtol=2e-5)
    g = g.cpu().numpy()
    np.testing.assert_allclose(g, gref, rtol=5e-4, atol=2e-4 * np.abs(gref).max())
    assert np.all(g[gref == 0] == 0)


@pytest.mark.parametrize("preset,d", [("nsf6", 10), ("nsf3", 4), ("maf3", 100)])
def test_default_spline_fit_runs_on_own_kernels_and_matches_reference_loop(preset, d, monkeypatch):
    """Flow.fit of the reference's default flow family never enters torch autograd: every optimiser step is the layer-wise
    kernels + fused clip / AdamW inside a CUDA graph, and the loss history follows the reference training loop
    (oracle/flow_ref.py: zuko module, torch AdamW, clip_grad_norm_) run on the same data with the same seed."""
    from pocomc_b200.flow import Flow, MaskedAutoregressiveFlow

    def boom(*a, **k):
        raise AssertionError("Flow.fit entered the autograd path")

    torch.manual_seed(31)
    x = torch.randn(700, d) * 1.4 + 0.3
    w = torch.rand(700) + 0.1
    torch.manual_seed(5)
    ref_flow = F.make_flow(d, preset)
    torch.manual_seed(5)
    f = Flow(d, preset)
    monkeypatch.setattr(MaskedAutoregressiveFlow, "forward_autograd", boom)
    torch.manual_seed(9)
    hist = f.fit(x, weights=w, epochs=6, batch_size=128, validation_split=0.8)
    assert len(hist["loss"]) == 6 and np.all(np.isfinite(hist["loss"])) and np.all(np.isfinite(hist["val_loss"]))
    assert hist["loss"][-1] < hist["loss"][0]
    torch.manual_seed(9)
    href = F.fit(ref_flow, x, weights=w, epochs=6, batch_size=128, validation_split=0.8)
    np.testing.assert_allclose(hist["loss"], href["loss"], rtol=5e-4)
    np.testing.assert_allclose(hist["val_loss"], href["val_loss"], rtol=5e-4)
