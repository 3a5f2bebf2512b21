"""CPU: the packed degree-sorted layout + single-sweep algorithm reproduce the zuko oracle's
1-pass forward and (D+1)-pass inverse."""
import numpy as np
import pytest
import torch

import flow_ref as F
from pocomc_b200 import made_layout as ML
from sweep_emul import pack, pack_stream, sweep, sweep_stream


def _raw(flow):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in flow.parameters()])


@pytest.mark.parametrize("preset,d", [("maf3", 2), ("maf3", 3), ("maf3", 4), ("maf6", 10), ("nsf3", 5),
                                      ("nsf6", 2), ("maf6", 32), ("nsf3", 13)])
def test_sweep_matches_oracle(preset, d, faithful_fp32_oracle):
    torch.manual_seed(d)
    flow = F.make_flow(d, preset)
    kind0 = F.PRESETS[preset][0]
    with torch.no_grad():
        for p in flow.parameters():          # make the (affine) flow far from identity
            p.mul_(1.5 if (d <= 5 and kind0 == "maf") else 1.0)
    kind, T = F.PRESETS[preset]
    lay = ML.build_layout(d, F.hidden_width(d), 3, T, ML.KIND_AFFINE if kind == "maf" else ML.KIND_RQS)
    raw = _raw(flow)
    assert raw.size == lay.raw_numel
    packed = pack(lay, raw)
    x = (torch.randn(40, d) * 1.7).float()
    with torch.no_grad():
        z, ladj = flow().transform.call_and_ladj(x)
        xi, li = flow().transform.inv.call_and_ladj(z)
    zs, ls = sweep(lay, packed, x.numpy(), inverse=False)
    np.testing.assert_allclose(zs, z.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(ls, ladj.numpy(), rtol=1e-4, atol=2e-5)
    xs, lis = sweep(lay, packed, z.numpy(), inverse=True)
    tol = 1e-4 if kind == "maf" else 5e-4
    np.testing.assert_allclose(xs, xi.numpy(), rtol=tol, atol=tol)
    np.testing.assert_allclose(lis, li.numpy(), rtol=tol, atol=tol)
    # the TMA-stream layout (v2 kernel) walks the same arithmetic in consumption order
    assert ML.stream_supported(d, F.hidden_width(d), 3, lay.kind)
    st = ML.build_stream(d, F.hidden_width(d), 3, T, lay.kind)
    sp = pack_stream(st, raw)
    assert np.all(st.chunks[:, 2] % 4 == 0) and np.all(st.chunks[:, 3] % 4 == 0) and st.chunks[:, 3].max() == st.slot_floats
    z2, l2 = sweep_stream(lay, st, sp, x.numpy(), inverse=False)
    np.testing.assert_allclose(z2, zs, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(l2, ls, rtol=1e-5, atol=1e-5)
    x2, li2 = sweep_stream(lay, st, sp, z.numpy(), inverse=True)
    tol2 = 5e-3 if (d <= 5 and kind == "maf") else tol     # the x1.5 weights amplify fp32 summation-order noise
    np.testing.assert_allclose(x2, xs, rtol=tol2, atol=tol2)
    np.testing.assert_allclose(li2, lis, rtol=tol2, atol=tol2)


def test_masks_match_oracle():
    for d, preset in ((4, "maf3"), (7, "nsf3")):
        flow = F.make_flow(d, preset)
        kind, T = F.PRESETS[preset]
        lay = ML.build_layout(d, F.hidden_width(d), 3, T, ML.KIND_AFFINE if kind == "maf" else ML.KIND_RQS)
        for t, tr in enumerate(flow.transform):
            ms = [m.mask.numpy() for m in tr.hyper.modules() if hasattr(m, "mask")]
            for a, b in zip(ms, ML.masks(lay, t)):
                np.testing.assert_array_equal(a, b)


def test_one_dim_rejected():
    with pytest.raises(ValueError):
        ML.build_layout(1, 32, 3, 3, ML.KIND_AFFINE)


@pytest.mark.parametrize("preset,d", [("maf3", 10), ("maf6", 32), ("maf3", 21), ("maf3", 16), ("maf3", 33), ("maf3", 36), ("maf3", 8),
                                      ("maf3", 25), ("maf3", 50), ("maf3", 12), ("maf3", 100),
                                      ("nsf3", 10), ("nsf6", 8), ("nsf3", 32), ("nsf3", 21), ("nsf3", 50), ("nsf3", 14)])
def test_block_triangular_layout_matches_oracle(preset, d):
    """made_layout.build_tri (image + tables of csrc/flow_tri.cu) walked by a numpy emulation of the kernel's schedule
    -- right-looking block updates with hi/lo TF32 operands, in-block fp32 substitution -- reproduces the oracle's
    1-pass forward and (D+1)-pass inverse."""
    from sweep_emul import pack_tri, sweep_tri
    torch.manual_seed(d)
    flow = F.make_flow(d, preset)
    with torch.no_grad():
        for p_ in flow.parameters():
            p_.mul_(1.3)
    raw = _raw(flow)
    T = int(preset[3:])
    from pocomc_b200 import tri_layout as TL
    kind = ML.KIND_AFFINE if preset.startswith("maf") else ML.KIND_RQS
    assert TL.tri_supported(d, F.hidden_width(d), 3, kind)
    tri = TL.build_tri(d, F.hidden_width(d), 3, T, kind)
    assert tri.smem_bytes <= TL.TRI_SMEM_BUDGET and tri.meta[TL.TRI_NCOLS] <= 512
    packed = pack_tri(tri, raw)
    x = torch.randn(37, d)
    with torch.no_grad():
        z, l = flow().transform.call_and_ladj(x)
        xi, li = flow().transform.inv.call_and_ladj(z)
    tol = 1.0 if kind == ML.KIND_AFFINE else 10.0          # the spline flows' bar is 5e-4
    zs, ls = sweep_tri(tri, packed, x.numpy(), inverse=False)
    np.testing.assert_allclose(zs, z.numpy(), rtol=2e-5 * tol, atol=2e-5 * tol)
    np.testing.assert_allclose(ls, l.numpy(), rtol=2e-5 * tol, atol=2e-5 * tol)
    xs, lis = sweep_tri(tri, packed, z.numpy(), inverse=True)
    np.testing.assert_allclose(xs, xi.numpy(), rtol=5e-5 * tol, atol=5e-5 * tol)
    np.testing.assert_allclose(lis, li.numpy(), rtol=5e-5 * tol, atol=5e-5 * tol)
    # plain TF32 (passes = 1) is visibly worse: the 3-pass split is what buys fp32 fidelity
    z1, _ = sweep_tri(tri, packed, x.numpy(), inverse=False, passes=1)
    assert np.abs(z1 - z.numpy()).max() > 10 * np.abs(zs - z.numpy()).max()


def test_block_triangular_support_matrix():
    from pocomc_b200 import tri_layout as TL
    sup = {d: TL.tri_supported(d, F.hidden_width(d), 3, ML.KIND_AFFINE) for d in (2, 6, 10, 21, 32, 50, 100, 200)}
    assert sup[10] and sup[21] and sup[32] and sup[50] and sup[100] and sup[200] and not sup[2] and not sup[6]
    # spline flows (zuko NSF): two blocks per window at most (256 output columns), so even 10-D uses the scratch area
    nsf = {d: TL.tri_supported(d, F.hidden_width(d), 3, ML.KIND_RQS) for d in (2, 10, 32, 50, 100, 200)}
    assert all(nsf[d] for d in (10, 32, 50, 100, 200)) and not nsf[2]
    assert TL.build_tri(10, 32, 3, 1, ML.KIND_RQS).meta[TL.TRI_NW] == 2
    # one window (no scratch area) up to 36 dimensions, several windows beyond
    assert TL.build_tri(32, 128, 3, 1, ML.KIND_AFFINE).ws_floats == 0
    big = TL.build_tri(200, 1024, 3, 1, ML.KIND_AFFINE)
    assert big.meta[TL.TRI_NW] >= 6 and big.ws_floats > 0 and big.smem_bytes <= TL.TRI_SMEM_BUDGET
