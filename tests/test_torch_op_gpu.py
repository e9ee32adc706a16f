"""GPU tests of torch.ops.cspn_b200.* (csrc/torch_op.cpp): same bits as the ctypes binding of the same C ABI, gradients
equal to the autograd.Function route, torch.library.opcheck, torch.compile(fullgraph=True) through the caller, CUDA-graph
capture."""
import pytest
import torch
import torch.nn as nn

import cspn_b200
from cspn_b200 import _lib, torch_op
from cspn_b200.cspn import _Propagate2dFn
from cspn_b200.synth import make_inputs, make_inputs_3d

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _loaded():
    torch_op.load()


@pytest.mark.parametrize('norm', ['8sum', '8sum_abs'])
@pytest.mark.parametrize('shape,n', [((2, 1, 40, 132), 24), ((1, 3, 13, 17), 5), ((4, 1, 228, 304), 24)])
def test_op_equals_ctypes_binding_bit_for_bit(shape, n, norm):
    B, C, H, W = shape
    g, d, s = [t.cuda() for t in make_inputs(3, B, C, H, W, 9, 'signed', 200)]
    a = torch.ops.cspn_b200.propagate2d(g, d, s, n, _lib.NORM2D[norm], _lib.ALGO_AUTO)
    b = cspn_b200.propagate2d(g, d, s, n, norm)
    assert torch.equal(a, b)
    assert torch.equal(torch.ops.cspn_b200.propagate2d(g, d, None, n, _lib.NORM2D[norm], 0), cspn_b200.propagate2d(g, d, None, n, norm))


def test_gradients_equal_the_autograd_function_route():
    g, d, s = [t.cuda() for t in make_inputs(5, 2, 1, 36, 64, 8, 'bernoulli', 80)]
    go = torch.rand_like(d)
    res = []
    for route in ('op', 'fn'):
        gc, dc = g.clone().requires_grad_(True), d.clone().requires_grad_(True)
        out = (torch.ops.cspn_b200.propagate2d(gc, dc, s, 12, 0, 0) if route == 'op'
               else _Propagate2dFn.apply(gc, dc, s, 12, '8sum', 0))
        out.backward(go)
        res.append((out.detach(), gc.grad, dc.grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # only one input needs a gradient
    gc = g.clone().requires_grad_(True)
    torch.ops.cspn_b200.propagate2d(gc, d, s, 12, 0, 0).backward(go)
    assert torch.equal(gc.grad, res[0][1])


def test_opcheck():
    g, d, s = [t.cuda() for t in make_inputs(7, 1, 1, 12, 16, 8, 'bernoulli', 10)]
    g.requires_grad_(True)
    d.requires_grad_(True)
    torch.library.opcheck(torch.ops.cspn_b200.propagate2d.default, (g, d, s, 6, 0, 0),
                          test_utils=('test_schema', 'test_faketensor', 'test_autograd_registration', 'test_aot_dispatch_dynamic'))
    g3, f3 = [t.cuda() for t in make_inputs_3d(1, 1, 1, 4, 6, 8)]
    torch.library.opcheck(torch.ops.cspn_b200.propagate3d.default, (g3.requires_grad_(True), f3.requires_grad_(True), 3, 1),
                          test_utils=('test_schema', 'test_faketensor', 'test_autograd_registration'))


class TinyCaller(nn.Module):
    """Stand-in for ResNet.forward's tail (torch_resnet_cspn_nyu.py:351,372-375)."""

    def __init__(self, step=8, norm='8sum'):
        super().__init__()
        self.gud = nn.Conv2d(4, 8, 3, padding=1, bias=False)
        self.dep = nn.Conv2d(4, 1, 3, padding=1, bias=False)
        self.post_process_layer = cspn_b200.Affinity_Propagate(step, 3, norm)

    def forward(self, x):
        sparse_depth = x.narrow(1, 3, 1).clone()
        return self.post_process_layer(self.gud(x), self.dep(x), sparse_depth)


def test_torch_compile_fullgraph_through_the_caller():
    torch.manual_seed(0)
    net = TinyCaller().cuda()
    x = torch.rand(2, 4, 32, 48, device='cuda')
    ref = net(x)
    ref.sum().backward()
    gref = net.gud.weight.grad.clone()
    net.zero_grad()
    cnet = torch.compile(net, fullgraph=True, backend='aot_eager')     # no graph break allowed; aot_eager keeps the test off inductor's codegen
    out = cnet(x)
    out.sum().backward()
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
    assert torch.allclose(net.gud.weight.grad, gref, rtol=1e-5, atol=1e-6)


def test_cuda_graph_capture_and_small_problem_latency():
    """cfg1's shape (1x1x228x304, N=24) is launch-latency territory: the op must be capturable (no sync, no host round trip
    between its launches), and a replay gives the latency floor of the kernel itself."""
    g, d, s = [t.cuda() for t in make_inputs(0, 1, 1, 228, 304)]
    ref = torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(50):
        graph.replay()
    ev[1].record()
    torch.cuda.synchronize()
    per_replay_us = ev[0].elapsed_time(ev[1]) * 1e3 / 50
    ev[0].record()
    for _ in range(50):
        torch.ops.cspn_b200.propagate2d(g, d, s, 24, 0, 0)
    ev[1].record()
    torch.cuda.synchronize()
    per_call_us = ev[0].elapsed_time(ev[1]) * 1e3 / 50
    print(f'cfg1 (1x1x228x304, N=24): graph replay {per_replay_us:.1f} us, eager op call {per_call_us:.1f} us')
    assert per_replay_us < 60
