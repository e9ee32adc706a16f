"""GPU test of the native adjoint (cspn2d_bwd_f32) against autograd through the reference's op sequence
(oracle/cspn_torch_port.py, bit-identical to cspn.py on the forward).  Called through the nn.Module, i.e. the way
train.py:196-199 uses the operator."""
import glob
import os

import numpy as np
import pytest
import torch

import cspn_b200
from cspn_b200.synth import make_inputs
from conftest import GOLDEN_DIR
from oracle import cspn_torch_port as tp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('norm', ['8sum', '8sum_abs'])
@pytest.mark.parametrize('shape,n,sparse', [((2, 1, 12, 16), 5, 'signed'), ((1, 2, 9, 20), 3, 'bernoulli'),
                                            ((1, 1, 24, 32), 24, 'bernoulli'), ((2, 1, 10, 15), 4, None)])
def test_gradients_match_reference_autograd(shape, n, sparse, norm):
    B, C, H, W = shape
    g, d, s = make_inputs(31 + H, B, C, H, W, 9, sparse, 25)
    go = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(1))
    # reference gradients: fp64 autograd through the reference's ops on CPU
    g64 = g.double().requires_grad_(True)
    d64 = d.double().requires_grad_(True)
    ref = tp.cspn2d_torch(g64, d64, None if s is None else s.double(), n, norm)
    ref.backward(go.double())
    # ours
    gc = g.cuda().requires_grad_(True)
    dc = d.cuda().requires_grad_(True)
    out = cspn_b200.Affinity_Propagate(n, 3, norm)(gc, dc, None if s is None else s.cuda())
    assert out.requires_grad
    out.backward(go.cuda())
    for ours, theirs, name in ((gc.grad, g64.grad, 'guidance'), (dc.grad, d64.grad, 'blur')):
        ours = ours.double().cpu()
        scale = theirs.abs().mean()
        err = (ours - theirs).abs()
        assert (err <= 2e-3 * (theirs.abs() + scale)).all(), (name, float(err.max()), float(scale))
    assert torch.count_nonzero(gc.grad[:, 8:]) == 0            # unused guidance channels get zero gradient


GRAD_GOLDENS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, 'grads', '*.npz')))


@pytest.mark.parametrize('name', GRAD_GOLDENS)
def test_gradients_match_the_reference_modules_own_autograd(name):
    """tests/golden/grads: gradients from autograd through the unmodified reference module (fp32, as train.py runs it)."""
    z = np.load(os.path.join(GOLDEN_DIR, 'grads', name + '.npz'))
    gc = torch.tensor(z['guidance']).cuda().requires_grad_(True)
    dc = torch.tensor(z['blur']).cuda().requires_grad_(True)
    s = torch.tensor(z['sparse_depth']).cuda() if 'sparse_depth' in z else None
    out = cspn_b200.Affinity_Propagate(int(z['prop_time']), 3, str(z['norm_type']))(gc, dc, s)
    out.backward(torch.tensor(z['grad_out']).cuda())
    for ours, ref, what in ((gc.grad, z['grad_guidance'], 'guidance'), (dc.grad, z['grad_blur'], 'blur')):
        ours = ours.cpu().numpy()
        scale = np.abs(ref).mean()
        err = np.abs(ours - ref)
        assert (err <= 3e-3 * (np.abs(ref) + scale)).all(), (what, float(err.max()), float(scale))


def test_training_step_through_the_module():
    """A miniature of train.py:196-199: conv -> CSPN -> loss -> backward -> SGD step changes the producer's weights."""
    torch.manual_seed(0)
    net = torch.nn.Conv2d(4, 9, 3, padding=1).cuda()            # produces 8 guidance channels + 1 blur depth
    post = cspn_b200.Affinity_Propagate(6, 3, '8sum_abs')
    opt = torch.optim.SGD(list(net.parameters()) + list(post.parameters()), lr=0.1)
    x = torch.rand(2, 4, 20, 24, device='cuda')
    target = torch.rand(2, 1, 20, 24, device='cuda')
    w0 = net.weight.detach().clone()
    losses = []
    for _ in range(3):
        y = net(x)
        out = post(y[:, :8], y[:, 8:9], x[:, 3:4].clone())
        loss = (out - target).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert not torch.equal(net.weight, w0) and all(l == l for l in losses)
