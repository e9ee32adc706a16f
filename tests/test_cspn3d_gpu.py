"""GPU parity of the 3D path against the (reference-unpinned) oracle, through the C ABI."""
import numpy as np
import pytest
import torch

import cspn_b200
from cspn_b200 import _lib
from cspn_b200.synth import make_inputs_3d
from oracle import c_oracle, cspn_numpy as onp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['26sum', '26sum_abs', 'paddle'])
@pytest.mark.parametrize('shape,n', [((2, 1, 6, 10, 12), 4), ((1, 2, 16, 24, 40), 12), ((1, 1, 1, 9, 11), 3),
                                      ((1, 1, 5, 1, 1), 2)])
def test_3d_against_oracle(shape, n, mode):
    B, C, D, H, W = shape
    g, f = make_inputs_3d(5, B, C, D, H, W, signed=(mode == '26sum'))
    ref = c_oracle.cspn3d(g.numpy(), f.numpy(), n, mode)
    out = cspn_b200.Affinity_Propagate3D(n, 3, mode)(g.cuda(), f.cuda()).cpu().numpy()
    ok, ratio, normwise = onp.parity_ok(out, ref, 1e-4)
    assert ok, (ratio, normwise)


def test_3d_full_size_properties():
    """BASELINE cfg4 shape (8,1,64,96,312), N=12: linearity, constant preservation, one slab vs the oracle."""
    B, C, D, H, W = 2, 1, 64, 96, 312          # two of the eight volumes: volumes are independent
    g, f = [t.cuda() for t in make_inputs_3d(6, B, C, D, H, W)]
    m = cspn_b200.Affinity_Propagate3D(12, 3, '26sum_abs')
    out = m(g, f)
    assert torch.isfinite(out).all()
    f2 = torch.rand_like(f)
    lhs = m(g, 0.5 * f + 3 * f2)
    rhs = 0.5 * out + 3 * m(g, f2)
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-4)
    const = torch.full_like(f, 1.25)
    assert torch.allclose(m(g, const), const, rtol=1e-5)
    assert torch.allclose(cspn_b200.Affinity_Propagate3D(12, 3, 'paddle')(g, const)[:, :, 12:-12, 12:-12, 12:-12],
                          const[:, :, 12:-12, 12:-12, 12:-12], rtol=1e-5)
    ref = c_oracle.cspn3d(g[:1].cpu().numpy(), f[:1].cpu().numpy(), 12, '26sum_abs')
    ok, ratio, normwise = onp.parity_ok(out[:1].cpu().numpy(), ref, 1e-4)
    assert ok, (ratio, normwise)


@pytest.mark.parametrize('cap_kb,launches', [(None, 1 + 4), (300, 2 * (1 + 4)), (100, 3 * (1 + 4))])
def test_3d_volume_groups_including_a_ragged_last_group(cap_kb, launches, monkeypatch):
    """One launch covers a group of volumes; the workspace cap decides the group size (3 volumes: 3, 2+1, 1+1+1)."""
    monkeypatch.setenv('CSPN_B200_3D_PADDLE', 'planes')              # the prep + weight-plane path (W % 4 != 0 fallback) is the one with groups
    if cap_kb:
        monkeypatch.setenv('CSPN_B200_MAX_WS3D_KB', str(cap_kb))     # one volume needs 29 * 6*10*16 * 4 B = 111 KB
    g, f = make_inputs_3d(4, 3, 2, 6, 10, 16)
    ref = c_oracle.cspn3d(g.numpy(), f.numpy(), 4, '26sum_abs')
    out = cspn_b200.propagate3d(g.cuda(), f.cuda(), 4, '26sum_abs')
    torch.cuda.synchronize()
    assert _lib.lib().cspn_last_launches() == launches
    ok, ratio, normwise = onp.parity_ok(out.cpu().numpy(), ref, 1e-4)
    assert ok, (ratio, normwise)


def test_3d_host_entry_point():
    g, f = make_inputs_3d(8, 3, 1, 8, 12, 16)
    a = cspn_b200.propagate3d(g, f, 5, 'paddle')
    b = cspn_b200.propagate3d(g.cuda(), f.cuda(), 5, 'paddle').cpu()
    assert torch.equal(a, b)


@pytest.mark.parametrize('mode', ['26sum', '26sum_abs', 'paddle'])
@pytest.mark.parametrize('shape,n', [((2, 1, 4, 6, 8), 3), ((1, 2, 5, 5, 12), 4)])
def test_3d_gradients_match_autograd_of_the_restatement(shape, n, mode):
    from oracle import cspn_torch_port as tp
    B, C, D, H, W = shape
    g, f = make_inputs_3d(9, B, C, D, H, W, signed=(mode != 'paddle'))
    g = g + 0.05 * torch.sign(g)                                   # keep |g| away from the kink of abs
    go = torch.randn(B, C, D, H, W, generator=torch.Generator().manual_seed(2))
    g64, f64 = g.double().requires_grad_(True), f.double().requires_grad_(True)
    ref = tp.cspn3d_torch(g64, f64, n, mode)
    assert np.allclose(ref.detach().numpy(), onp.cspn3d(g.numpy(), f.numpy(), n, mode, dtype=np.float64), rtol=1e-9, atol=1e-9)
    ref.backward(go.double())
    gc, fc = g.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    out = cspn_b200.Affinity_Propagate3D(n, 3, mode)(gc, fc)
    out.backward(go.cuda())
    for ours, theirs, name in ((gc.grad, g64.grad, 'guidance'), (fc.grad, f64.grad, 'feat')):
        ours = ours.double().cpu()
        scale = theirs.abs().mean()
        err = (ours - theirs).abs()
        assert (err <= 2e-3 * (theirs.abs() + scale)).all(), (name, mode, float(err.max()), float(scale))


@pytest.mark.parametrize('mode', ['paddle', '26sum_abs'])
def test_3d_per_channel_gates_like_demo_py(mode):
    """cspn_paddle/demo.py:28-45: with C > 1 feature channels the guide carries 26*C channels and channel c is propagated
    with its own slice [26c, 26c+26).  Checked against the oracle applied slice by slice, forward and backward."""
    B, C, D, H, W, n = 2, 3, 5, 8, 12, 4
    gen = torch.Generator().manual_seed(11)
    guide = torch.rand(B, 26 * C, D, H, W, generator=gen)
    feat = torch.rand(B, C, D, H, W, generator=gen)
    layer = cspn_b200.Affinity_Propagate3D(n, 3, mode)
    gc, fc = guide.cuda().requires_grad_(True), feat.cuda().requires_grad_(True)
    out = layer(gc, fc)
    assert out.shape == feat.shape
    ref = np.stack([c_oracle.cspn3d(guide[:, 26 * c:26 * (c + 1)].contiguous().numpy(), feat[:, c:c + 1].contiguous().numpy(), n, mode)[:, 0]
                    for c in range(C)], 1)
    ok, ratio, normwise = onp.parity_ok(out.detach().cpu().numpy(), ref, 1e-4)
    assert ok, (ratio, normwise)
    out.sum().backward()
    # gradients: the same thing done channel by channel with shared-gate calls
    gg = torch.zeros_like(guide)
    gf = torch.zeros_like(feat)
    for c in range(C):
        g1 = guide[:, 26 * c:26 * (c + 1)].contiguous().cuda().requires_grad_(True)
        f1 = feat[:, c:c + 1].contiguous().cuda().requires_grad_(True)
        layer(g1, f1).sum().backward()
        gg[:, 26 * c:26 * (c + 1)] = g1.grad.cpu()
        gf[:, c:c + 1] = f1.grad.cpu()
    assert torch.allclose(gc.grad.cpu(), gg, rtol=1e-5, atol=1e-7) and torch.allclose(fc.grad.cpu(), gf, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('mode', ['26sum', '26sum_abs', 'paddle'])
def test_3d_direct_path_agrees_with_the_weight_plane_path(mode, monkeypatch):
    """W % 4 == 0: gates straight from the raw guidance, one launch per step, no weight planes; CSPN_B200_3D_PADDLE=planes
    forces prep + 27 weight planes (the path odd widths take).  Same arithmetic up to one rounding per weight."""
    g, f = [t.cuda() for t in make_inputs_3d(9, 2, 2, 7, 11, 24, signed=(mode == '26sum'))]
    direct = cspn_b200.propagate3d(g, f, 6, mode)
    torch.cuda.synchronize()
    assert _lib.lib().cspn_last_launches() == 6
    monkeypatch.setenv('CSPN_B200_3D_PADDLE', 'planes')
    planes = cspn_b200.propagate3d(g, f, 6, mode)
    torch.cuda.synchronize()
    assert _lib.lib().cspn_last_launches() == 7
    ok, ratio, normwise = onp.parity_ok(direct.cpu().numpy(), planes.cpu().numpy(), 1e-5)
    assert ok, (ratio, normwise)


def test_3d_degenerate_gates_follow_ieee_division():
    """A voxel whose gates are all zero (own location for 'paddle', the 26 gathered taps for '26sum_abs'): 0 / 0 = NaN, and the
    NaN front then moves one voxel per step -- on the direct path exactly as in the oracle."""
    g, f = make_inputs_3d(3, 1, 1, 6, 8, 12)
    g[:, :, 1:4, 2:5, 3:6] = 0.0
    for mode in ('26sum_abs', 'paddle'):
        for n in (1, 2):
            ref = c_oracle.cspn3d(g.numpy(), f.numpy(), n, mode)
            out = cspn_b200.propagate3d(g.cuda(), f.cuda(), n, mode).cpu().numpy()
            assert np.isnan(ref).any()
            assert np.array_equal(np.isnan(out), np.isnan(ref)), (mode, n)
            ok, ratio, normwise = onp.parity_ok(out, ref, 1e-4)
            assert ok, (mode, n, ratio, normwise)
