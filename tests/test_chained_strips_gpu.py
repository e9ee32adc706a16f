"""Chained strips of the 2D cluster kernel: the strips of an image are processed left to right and each
hands the column left of its right neighbour on, step by step, through a history block in the workspace.  A pixel's
arithmetic does not depend on the tiling, so the chained and the unchained plan must agree bit for bit whenever they use the
same thread patch; both are also held against the C oracle.  CSPN_B200_CHAIN=1 / 0 forces / forbids chaining (read per call)."""
import os

import numpy as np
import pytest
import torch

import cspn_b200
from cspn_b200 import _lib
from cspn_b200.synth import make_inputs
from oracle import c_oracle, cspn_numpy as onp

pytestmark = pytest.mark.gpu


def run(g, d, s, n, norm, chain):
    old = os.environ.get('CSPN_B200_CHAIN')
    try:
        if chain is None:
            os.environ.pop('CSPN_B200_CHAIN', None)
        else:
            os.environ['CSPN_B200_CHAIN'] = '1' if chain else '0'
        B, C, H, W = d.shape
        plan = cspn_b200.describe_plan(B, C, H, W, n, _lib.ALGO_CLUSTER)
        out = cspn_b200.propagate2d(g, d, s, n, norm, _lib.ALGO_CLUSTER)
        torch.cuda.synchronize()
        return out, plan
    finally:
        if old is None:
            os.environ.pop('CSPN_B200_CHAIN', None)
        else:
            os.environ['CSPN_B200_CHAIN'] = old


def patch_of(plan):
    return plan.split('patch ')[1].split(' px/thread')[0]


@pytest.mark.parametrize('B,C,H,W,n,norm,with_sparse', [
    (2, 1, 40, 300, 10, '8sum', True),          # three strips, one CTA
    (3, 2, 100, 520, 24, '8sum', True),         # five strips, clusters of several CTAs, C > 1
    (2, 1, 228, 304, 24, '8sum_abs', False),    # NYU shape
    (4, 1, 352, 1216, 24, '8sum', True),        # KITTI shape: 12 chained strips instead of 15
    (2, 1, 64, 304, 48, '8sum', True),          # several passes, each chained
    (1, 1, 33, 260, 31, '8sum', True),          # odd height, 31 steps (halo rounded up to 32)
    (5, 1, 24, 1028, 7, '8sum', False),         # many short strips
])
def test_chained_plan_matches_unchained_plan_and_oracle(B, C, H, W, n, norm, with_sparse):
    g, d, s = make_inputs(7 * H + W + n, B, C, H, W)
    if not with_sparse:
        s = None
    gc, dc, sc = g.cuda(), d.cuda(), None if s is None else s.cuda()
    plain, plan_u = run(gc, dc, sc, n, norm, chain=False)
    chained, plan_c = run(gc, dc, sc, n, norm, chain=True)
    assert 'chained' in plan_c and 'chained' not in plan_u, (plan_c, plan_u)
    if patch_of(plan_c) == patch_of(plan_u):
        assert torch.equal(chained, plain), f'max diff {(chained - plain).abs().max().item():.3g}\n{plan_c}\n{plan_u}'
    else:
        torch.testing.assert_close(chained, plain, rtol=2e-6, atol=2e-6)
    ref = c_oracle.cspn2d(g.numpy(), d.numpy(), None if s is None else s.numpy(), n, norm)
    ok, ratio, normwise = onp.parity_ok(chained.cpu().numpy(), ref, 1e-4)
    assert ok, f'violation ratio {ratio:.3g}, normwise {normwise:.3g}'


def test_large_batch_chains_by_default_and_small_batch_does_not():
    """The planner chains when every cluster has at least two rounds of tasks (a task's left neighbour is then long done)."""
    small = cspn_b200.describe_plan(2, 1, 228, 304, 24, _lib.ALGO_CLUSTER)
    large = cspn_b200.describe_plan(64, 1, 228, 304, 24, _lib.ALGO_CLUSTER)
    assert 'chained' not in small, small
    assert 'chained' in large, large
    g, d, s = make_inputs(5, 8, 1, 228, 304)
    g, d, s = [t.repeat(8, 1, 1, 1).contiguous().cuda() for t in (g, d, s)]
    auto, plan = run(g, d, s, 24, '8sum', chain=None)
    assert 'chained' in plan
    plain, _ = run(g, d, s, 24, '8sum', chain=False)
    if patch_of(plan) == patch_of(_):
        assert torch.equal(auto, plain)
    else:
        torch.testing.assert_close(auto, plain, rtol=2e-6, atol=2e-6)
    # replicas of the same 8 images must come out identical whatever task / cluster computed them
    assert torch.equal(auto[:8], auto[8:16]) and torch.equal(auto[:8], auto[56:64])


def test_chained_call_is_repeatable_and_graph_capturable():
    """The flags are reset by a memset node in front of the launch: replays must not see stale 'ready' words."""
    g, d, s = [t.cuda() for t in make_inputs(3, 3, 1, 80, 520)]
    os.environ['CSPN_B200_CHAIN'] = '1'
    try:
        first = cspn_b200.propagate2d(g, d, s, 12, '8sum', _lib.ALGO_CLUSTER).clone()
        out = torch.empty_like(d)
        cspn_b200.propagate2d(g, d, s, 12, '8sum', _lib.ALGO_CLUSTER, out=out)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            m = cspn_b200.Affinity_Propagate(12, 3, '8sum')
            res = m(g, d, s)
        for _ in range(3):
            res.zero_()
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(res, first)
    finally:
        os.environ.pop('CSPN_B200_CHAIN', None)


@pytest.mark.parametrize('group', [1, 2, 3])
def test_task_groups_do_not_change_the_result(group):
    """CSPN_B200_CHAIN_GROUP: tasks run group by group (strip-major inside a group of image-channels; the last group takes the
    remainder), which only reorders them."""
    g, d, s = [t.cuda() for t in make_inputs(11, 7, 1, 48, 400)]
    plain, _ = run(g, d, s, 9, '8sum', chain=False)
    os.environ['CSPN_B200_CHAIN_GROUP'] = str(group)
    try:
        chained, plan = run(g, d, s, 9, '8sum', chain=True)
    finally:
        os.environ.pop('CSPN_B200_CHAIN_GROUP', None)
    assert 'chained' in plan
    assert torch.equal(chained, plain) if patch_of(plan) == patch_of(_) else torch.allclose(chained, plain, rtol=2e-6, atol=2e-6)
