"""Edge cases of the cluster kernel's tiling (strips, halos, bands) and of the AUTO fallback, through the C ABI."""
import numpy as np
import pytest
import torch

import cspn_b200
from cspn_b200 import _lib
from cspn_b200.synth import make_inputs
from oracle import c_oracle, cspn_numpy as onp

pytestmark = pytest.mark.gpu


def check(B, C, H, W, n, norm='8sum', gch=8, sparse='signed', algo=_lib.ALGO_AUTO, expect=None):
    g, d, s = make_inputs(H * 7 + W + n, B, C, H, W, gch, sparse, 300)
    ref = c_oracle.cspn2d(g.numpy(), d.numpy(), None if s is None else s.numpy(), n, norm)
    out = cspn_b200.propagate2d(g.cuda(), d.cuda(), None if s is None else s.cuda(), n, norm, algo)
    torch.cuda.synchronize()
    used = _lib.ALGO_NAMES[_lib.lib().cspn_last_algo()]
    if expect:
        assert used == expect, (used, cspn_b200.describe_plan(B, C, H, W, n, algo))
    ok, ratio, normwise = onp.parity_ok(out.cpu().numpy(), ref, 1e-4)
    assert ok, (B, C, H, W, n, norm, used, ratio, normwise)


@pytest.mark.parametrize('H,W,n', [
    (40, 128, 24),      # exactly one CTA band, one strip
    (41, 132, 24),      # one row / one float4 more than a tile
    (352, 1216, 48),    # halo 48: many narrow strips
    (228, 912, 24),     # the reference's own KITTI crop (kitti_dataset_loader.py:85)
    (64, 2048, 12),     # wide and flat
    (600, 64, 10),      # tall and thin: 15 bands, a single half-empty strip
    (3, 4, 2), (1, 128, 5), (16, 260, 30),
])
def test_tilings_match_oracle(H, W, n):
    check(1, 1, H, W, n, expect='cluster')


@pytest.mark.parametrize('H,W,n', [
    (700, 64, 4),       # taller than 16 CTAs of 40 rows: overlapping row bands
    (1000, 256, 30),    # bands x strips, both with halos
    (650, 128, 3),      # one strip, bands with a 3-row halo
    (24, 512, 70),      # halo wider than a strip can carry: several passes through HBM
    (228, 304, 48),     # cfg3's longest run: 3 passes of 16 steps
    (100, 132, 131),    # odd step count split into passes of unequal length
    (352, 1216, 49),
])
def test_tall_images_and_long_runs_stay_on_the_cluster_path(H, W, n):
    check(1, 1, H, W, n, expect='cluster')
    assert _lib.lib().cspn_last_launches() >= 1


def test_multi_pass_with_channels_abs_norm_and_no_sparse():
    check(2, 3, 60, 260, 90, '8sum_abs', gch=9, sparse=None, expect='cluster')
    assert _lib.lib().cspn_last_launches() > 1


def test_multi_pass_through_the_host_entry_point():
    g, d, s = make_inputs(11, 3, 1, 64, 256, 8, 'signed', 100)
    ref = c_oracle.cspn2d(g.numpy(), d.numpy(), s.numpy(), 80, '8sum')
    out = cspn_b200.propagate2d(g, d, s, 80, '8sum')        # CPU tensors -> cspn2d_fwd_f32_host
    assert not out.is_cuda and onp.parity_ok(out.numpy(), ref, 1e-4)[0]


def test_missing_workspace_is_reported_not_ignored():
    g, d, s = [t.cuda() for t in make_inputs(3, 1, 1, 64, 256)]
    L = _lib.lib()
    out = torch.empty_like(d)
    need = L.cspn2d_workspace_bytes(1, 1, 64, 256, 120, _lib.ALGO_CLUSTER)
    assert need == d.numel() * 4
    rc = L.cspn2d_fwd_f32(g.data_ptr(), d.data_ptr(), s.data_ptr(), out.data_ptr(), 1, 1, 64, 256, 8, 120, 0,
                          _lib.ALGO_CLUSTER, None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == -2 and b'workspace' in L.cspn_last_error()


def test_shapes_the_cluster_kernel_declines_fall_back_to_generic():
    check(1, 1, 20, 18, 4, expect='generic')        # W % 4 != 0
    with pytest.raises(cspn_b200.CspnError):
        check(1, 1, 20, 18, 4, algo=_lib.ALGO_CLUSTER)


def test_channels_and_extra_guidance_on_the_cluster_path():
    check(3, 4, 100, 256, 12, '8sum_abs', gch=10, expect='cluster')
    check(2, 2, 60, 384, 8, '8sum', gch=8, sparse=None, expect='cluster')


def test_misaligned_views_are_handled():
    g, d, s = make_inputs(5, 2, 1, 48, 132, 9, 'bernoulli', 50)
    gc = g.cuda()[:, 1:]            # channel-offset view: contiguous() is taken by the wrapper, base stays 16B aligned
    ref = c_oracle.cspn2d(g[:, 1:].numpy(), d.numpy(), s.numpy(), 6, '8sum')
    out = cspn_b200.propagate2d(gc, d.cuda(), s.cuda(), 6, '8sum')
    assert onp.parity_ok(out.cpu().numpy(), ref, 1e-4)[0]


def test_tensors_at_odd_storage_offsets_take_the_generic_path():
    """A contiguous view whose base is only 4-byte aligned cannot feed TMA / float4 accesses: AUTO must not fail."""
    g, d, s = make_inputs(7, 1, 1, 40, 128)
    ref = c_oracle.cspn2d(g.numpy(), d.numpy(), s.numpy(), 6, '8sum')
    flat = torch.empty(d.numel() + 1, device='cuda')
    dv = flat[1:].view(d.shape)
    dv.copy_(d)
    assert dv.data_ptr() % 16 == 4 and dv.is_contiguous()
    out = cspn_b200.propagate2d(g.cuda(), dv, s.cuda(), 6, '8sum')
    torch.cuda.synchronize()
    assert _lib.ALGO_NAMES[_lib.lib().cspn_last_algo()] == 'generic'
    assert onp.parity_ok(out.cpu().numpy(), ref, 1e-4)[0]


@pytest.mark.parametrize('n', [24, 48])      # 48: three passes, i.e. three launches inside the graph
def test_cuda_graph_capture_and_replay(n):
    """No allocation / synchronisation inside the C ABI call: it can be captured in a CUDA graph."""
    g, d, s = [t.cuda() for t in make_inputs(3, 4, 1, 228, 304)]
    L = _lib.lib()
    out = torch.empty_like(d)
    ws_bytes = L.cspn2d_workspace_bytes(4, 1, 228, 304, n, _lib.ALGO_AUTO)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device='cuda')
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        def call():
            rc = L.cspn2d_fwd_f32(g.data_ptr(), d.data_ptr(), s.data_ptr(), out.data_ptr(), 4, 1, 228, 304, 8, n, 0,
                                  _lib.ALGO_AUTO, ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, L.cspn_last_error()
        call()                                       # warm-up outside capture (sets function attributes)
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(graph, stream=st):
            call()
    ref = cspn_b200.propagate2d(g, d, s, n, '8sum')
    d.mul_(0.5)                                      # new input values, same buffers
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(out, 0.5 * ref, rtol=1e-5, atol=1e-6)   # the map is linear in blur_depth
