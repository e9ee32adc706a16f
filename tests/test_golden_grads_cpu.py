"""Backward parity is pinned to the reference too: tests/golden/grads/*.npz hold gradients produced by autograd through
the unmodified reference module (tests/golden/make_golden_grads.py).  Here: the torch-op port (the checker of the GPU
backward tests) and the numpy model of the staged cluster backward must reproduce them."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import cspn_torch_port as tp
from test_staged_backward_model_cpu import staged_backward

NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, 'grads', '*.npz')))


def close(ours, ref, tol):
    scale = np.abs(ref).mean()
    return (np.abs(ours - ref) <= tol * (np.abs(ref) + scale)).all()


def test_gradient_goldens_exist():
    assert len(NAMES) >= 6


@pytest.mark.parametrize('name', NAMES)
def test_port_autograd_reproduces_the_reference_gradients(name):
    z = np.load(os.path.join(GOLDEN_DIR, 'grads', name + '.npz'))
    g = torch.tensor(z['guidance']).double().requires_grad_(True)
    d = torch.tensor(z['blur']).double().requires_grad_(True)
    s = torch.tensor(z['sparse_depth']).double() if 'sparse_depth' in z else None
    out = tp.cspn2d_torch(g, d, s, int(z['prop_time']), str(z['norm_type']))
    out.backward(torch.tensor(z['grad_out']).double())
    assert close(out.detach().numpy(), z['out'], 1e-4)
    assert close(g.grad.numpy(), z['grad_guidance'], 1e-4)       # fp32 reference vs fp64 port: measured <= 1.5e-5
    assert close(d.grad.numpy(), z['grad_blur'], 1e-4)


@pytest.mark.parametrize('name', NAMES)
def test_adjoint_formulation_reproduces_the_reference_gradients(name):
    """The hand-derived adjoint (lambda recursion, Gw/Gc gather, chain rule through the normalisation) that both native
    backward paths implement, in numpy fp64."""
    z = np.load(os.path.join(GOLDEN_DIR, 'grads', name + '.npz'))
    s = z['sparse_depth'].astype(np.float64) if 'sparse_depth' in z else None
    gg, gd = staged_backward(z['guidance'].astype(np.float64), z['blur'].astype(np.float64), s,
                             z['grad_out'].astype(np.float64), [int(z['prop_time'])], str(z['norm_type']))
    assert close(gg, z['grad_guidance'], 1e-4)
    assert close(gd, z['grad_blur'], 1e-4)
