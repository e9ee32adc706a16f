"""The drop-in, exercised with the reference's own caller (only where /root/reference is mounted, i.e. the build
container): `dropin/cspn.py` first on sys.path makes torch_resnet_cspn_nyu.py:12 (`import cspn as post_process`)
resolve to cspn_b200, the model builds our module from its cspn_config (:344-347), and its forward hands the module
exactly the tensors INTEGRATION.md promises (:351, :372-375).  The propagation itself needs a GPU and is replaced by a
recorder here."""
import importlib
import os
import sys

import pytest
import torch

import cspn_b200

REF_MODELS = '/root/reference/cspn_pytorch/models'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason='reference tree not mounted')


@pytest.fixture()
def reference_model_module(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, 'dropin'))     # before ./models, as INTEGRATION.md option A says
    monkeypatch.setattr(sys, 'path', sys.path + [REF_MODELS])
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)   # Unpool's ctor calls .cuda() (:50)
    for name in ('cspn', 'torch_resnet_cspn_nyu'):
        sys.modules.pop(name, None)
    mod = importlib.import_module('torch_resnet_cspn_nyu')
    yield mod
    for name in ('cspn', 'torch_resnet_cspn_nyu'):
        sys.modules.pop(name, None)


def test_reference_model_builds_and_calls_our_module(reference_model_module):
    m = reference_model_module
    assert m.post_process.Affinity_Propagate is cspn_b200.Affinity_Propagate
    net = m.resnet50(pretrained=False)        # the only depth the reference decoder fits (2048-ch bottleneck)
    layer = net.post_process_layer
    assert isinstance(layer, cspn_b200.Affinity_Propagate)
    assert (layer.prop_time, layer.prop_kernel, layer.norm_type) == (24, 3, '8sum')
    assert list(layer.parameters()) == [] and list(layer.buffers()) == []   # optim.SGD(net.parameters()) unchanged

    seen = {}

    def recorder(guidance, blur_depth, sparse_depth=None):
        seen.update(guidance=guidance, blur=blur_depth, sparse=sparse_depth)
        return blur_depth

    layer.forward = recorder
    x = torch.rand(2, 4, 228, 304)
    x[:, 3] *= (torch.rand(2, 228, 304) < 0.01)                 # sparse depth channel (nyu_dataset_loader.py:141-143)
    net.eval()
    with torch.no_grad():
        out = net(x)
    g, d, s = seen['guidance'], seen['blur'], seen['sparse']
    assert g.shape == (2, 8, 228, 304) and d.shape == (2, 1, 228, 304) and s.shape == (2, 1, 228, 304)
    assert g.dtype == d.dtype == s.dtype == torch.float32
    assert g.is_contiguous() and d.is_contiguous() and s.is_contiguous()
    assert torch.equal(s, x[:, 3:4])                              # the detached clone of the input's depth channel
    assert (g < 0).any()                                          # guidance is a plain conv output: signed
    assert out.shape == (2, 1, 228, 304)
    cspn_b200.cspn._check_inputs_2d(g, d, s)                      # what our forward would accept
