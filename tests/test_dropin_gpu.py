"""The module behaves like the reference's inside its callers: DataParallel replicas (eval.py:117), no_grad
evaluation (train.py:247), checkpoints with the reference's stray sum_conv key, prop_time sweep."""
import pytest
import torch
import torch.nn as nn

import cspn_b200
from cspn_b200.synth import make_inputs
from oracle import cspn_numpy as onp

pytestmark = pytest.mark.gpu


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


def test_caller_matches_oracle_and_loads_reference_style_checkpoint():
    torch.manual_seed(0)
    net = TinyCaller().cuda().eval()
    sd = {'module.' + k: v for k, v in net.state_dict().items()}
    sd['module.post_process_layer.sum_conv.weight'] = torch.ones(1, 8, 1, 1, 1)      # what reference checkpoints carry
    stripped = {k[len('module.'):]: v for k, v in sd.items()}                        # update_model.remove_moudle
    known = {k: v for k, v in stripped.items() if k in net.state_dict()}             # update_model.update_model
    net.load_state_dict(known)
    x = torch.rand(2, 4, 30, 40).cuda()
    with torch.no_grad():
        out = net(x)
        ref = onp.cspn2d(net.gud(x).cpu().numpy(), net.dep(x).cpu().numpy(), x[:, 3:4].cpu().numpy(), 8, '8sum')
    ok, ratio, normwise = onp.parity_ok(out.cpu().numpy(), ref, 1e-4)
    assert ok, (ratio, normwise)


def test_dataparallel_wrapper_like_eval_py():
    net = nn.DataParallel(TinyCaller().cuda(), device_ids=list(range(torch.cuda.device_count()))).eval()
    x = torch.rand(4, 4, 24, 32).cuda()
    with torch.no_grad():
        a = net(x)
        b = net.module(x)
    assert torch.equal(a.cpu(), b.cpu())


@pytest.mark.parametrize('n', [0, 1, 2, 3, 4, 8, 16, 24, 48])
def test_iteration_sweep_both_paths_agree(n):
    g, d, s = [t.cuda() for t in make_inputs(n, 2, 1, 64, 128)]
    a = cspn_b200.propagate2d(g, d, s, n, '8sum', cspn_b200.ALGO_GENERIC)
    b = cspn_b200.propagate2d(g, d, s, n, '8sum', cspn_b200.ALGO_AUTO)
    if n == 0:
        assert a is d and b is d
        return
    ok, ratio, normwise = onp.parity_ok(b.cpu().numpy(), a.cpu().numpy(), 1e-4)
    assert ok and normwise < 1e-5, (ratio, normwise)
