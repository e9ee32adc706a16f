"""CPU tests of the dispatcher registration (csrc/torch_op.cpp + cspn_b200/torch_op.py): the ops exist with the documented
schema, have no CPU kernel (no silent fallback), and their fake kernels / autograd formulas let PyTorch trace the call the
reference model makes (torch_resnet_cspn_nyu.py:375) without touching a GPU."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode
from torch.fx.experimental.proxy_tensor import make_fx

import cspn_b200
from cspn_b200 import build, torch_op


@pytest.fixture(scope='module', autouse=True)
def _built():
    build.build_torch_op()
    torch_op.load()


def test_ops_are_registered_with_the_documented_schema():
    s = str(torch.ops.cspn_b200.propagate2d.default._schema)
    assert 'Tensor guidance, Tensor blur_depth, Tensor? sparse_depth, int prop_time, int norm_type, int algo' in s
    for name in ('propagate2d', 'propagate2d_backward', 'propagate3d', 'propagate3d_backward'):
        assert hasattr(torch.ops.cspn_b200, name)


def test_no_cpu_kernel_exists():
    with pytest.raises(NotImplementedError, match='CPU'):
        torch.ops.cspn_b200.propagate2d(torch.zeros(1, 8, 4, 4), torch.zeros(1, 1, 4, 4), None, 2, 0, 0)
    with pytest.raises(NotImplementedError, match='CPU'):
        torch.ops.cspn_b200.propagate3d(torch.zeros(1, 26, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4), 2, 1)


def test_fake_kernels_propagate_shapes_on_a_box_without_a_gpu():
    with FakeTensorMode():
        g = torch.empty(2, 12, 16, 32, device='cuda')
        d = torch.empty(2, 3, 16, 32, device='cuda')
        s = torch.empty(2, 1, 16, 32, device='cuda')
        out = torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
        assert out.shape == d.shape and out.device == d.device and out.dtype == torch.float32
        gg, gd = torch.ops.cspn_b200.propagate2d_backward(g, d, s, out, 24, 0, True, False)
        assert gg.shape == g.shape and gd.numel() == 0
        g3, f3 = torch.empty(1, 26, 4, 8, 12, device='cuda'), torch.empty(1, 2, 4, 8, 12, device='cuda')
        assert torch.ops.cspn_b200.propagate3d(g3, f3, 12, 2).shape == f3.shape


def test_forward_and_backward_trace_to_one_node_each():
    """make_fx over forward + autograd.grad on meta tensors (the autograd engine wants a CUDA context for cuda fakes, and
    this box has none): the graph holds exactly the two registered ops, i.e. the backward is traceable as well."""
    def f(g, d, s):
        out = torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
        gg, gd = torch.autograd.grad(out.sum(), (g, d))
        return out, gg, gd

    g = torch.empty(1, 8, 12, 16, device='meta', requires_grad=True)
    d = torch.empty(1, 1, 12, 16, device='meta', requires_grad=True)
    s = torch.empty(1, 1, 12, 16, device='meta')
    gm = make_fx(f)(g, d, s)
    targets = [str(n.target) for n in gm.graph.nodes if n.op == 'call_function']
    assert sum('cspn_b200.propagate2d.default' in t for t in targets) == 1
    assert sum('cspn_b200.propagate2d_backward.default' in t for t in targets) == 1


def test_module_takes_the_dispatcher_route_when_the_shim_is_built():
    m = cspn_b200.Affinity_Propagate(24, 3)
    assert m._use_op is True
    with FakeTensorMode():
        out = m(torch.empty(1, 8, 8, 8, device='cuda'), torch.empty(1, 1, 8, 8, device='cuda'), torch.empty(1, 1, 8, 8, device='cuda'))
        assert out.shape == (1, 1, 8, 8)
