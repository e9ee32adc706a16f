"""Load the UNMODIFIED reference module when /root/reference is present.  TEST INFRASTRUCTURE ONLY.

The reference's forward calls `.cuda()` on a freshly built ones-weight
(/root/reference/cspn_pytorch/models/cspn.py:50), so it cannot run on a CPU-only
host as is.  `load_reference()` imports the file by path and, only while a reference
forward runs on CPU tensors, makes `Tensor.cuda` the identity.  Nothing is copied.

/root/reference does not exist on the GPU box: callers must check `available()`.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os

REF_FILE = '/root/reference/cspn_pytorch/models/cspn.py'


def available() -> bool:
    return os.path.isfile(REF_FILE)


def load_reference():
    spec = importlib.util.spec_from_file_location('_reference_cspn', REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def cuda_identity_shim():
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def reference_forward(guidance, blur_depth, sparse_depth, prop_time, norm_type):
    """Run reference Affinity_Propagate on CPU torch tensors; returns a torch tensor."""
    import torch
    mod = load_reference()
    layer = mod.Affinity_Propagate(prop_time, 3, norm_type)
    with torch.no_grad(), cuda_identity_shim():
        return layer(guidance, blur_depth, sparse_depth)


def reference_gradients(guidance, blur_depth, sparse_depth, grad_out, prop_time, norm_type):
    """autograd through the reference module itself (what train.py:196-199 differentiates): returns
    (out, d out / d guidance . grad_out, d out / d blur_depth . grad_out) as torch tensors."""
    mod = load_reference()
    layer = mod.Affinity_Propagate(prop_time, 3, norm_type)
    g = guidance.clone().requires_grad_(True)
    d = blur_depth.clone().requires_grad_(True)
    with cuda_identity_shim():
        out = layer(g, d, sparse_depth)
        out.backward(grad_out)
    return out.detach(), g.grad, d.grad
