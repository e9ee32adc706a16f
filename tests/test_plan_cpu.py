"""The cluster path's planner, checked on the CPU: geometry invariants of every plan, and a replay of the plan
(passes x bands x strips, each tile iterated in isolation with zero boundaries, useful region kept) on the numpy
oracle, which must reproduce the untiled result.  This is the halo argument of cspn2d_cluster.cu, executed."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from cspn_b200 import _lib
from oracle import cspn_numpy as onp


def check_axis(tiles, L, T, halo, align, chained=False):
    """tiles = [[t0, u0, u1], ...] must partition [0, L) into useful ranges, each inside its tile minus the halo.
    Chained tiles get their left neighbour column handed over: their left edge is exact and starts the useful range."""
    assert tiles[0][1] == 0 and tiles[-1][2] == L
    for i, (t0, u0, u1) in enumerate(tiles):
        assert t0 % align == 0 and u0 % align == 0 and t0 >= 0
        assert u0 < u1
        if i:
            assert u0 == tiles[i - 1][2]                       # no gap, no overlap of stored ranges
        assert t0 <= u0 and u1 <= t0 + T                       # stored range inside the tile
        if chained:
            assert u0 == t0
        elif t0 > 0:
            assert u0 - t0 >= halo                             # a cut side keeps >= halo stale positions out
        if t0 + T < L:
            assert t0 + T - u1 >= halo


def check_plan(H, W, n, chained=False):
    info = _lib.plan_info(H, W, n, chained)
    if chained and not info['supported']:
        return info
    assert info['supported'], info
    assert info['chained'] == chained
    assert sum(p['count'] * p['iters'] for p in info['passes']) == n
    for p in info['passes']:
        assert 1 <= p['cs'] <= 16 and p['RB'] == p['PR'] * p['NW'] and p['TW'] == 128
        assert len(p['strips']) <= 128 and len(p['bands']) <= 64
        check_axis(p['strips'], W, p['TW'], p['iters'], 4, chained)
        check_axis(p['bands'], H, p['cs'] * p['RB'], p['iters'], 1)
        if chained:
            assert len(p['bands']) == 1 and len(p['strips']) >= 2 and p['iters'] <= 32 and p['NW'] == 8
    return info


@pytest.mark.parametrize('H,W,n', [(352, 1216, 24), (228, 304, 48), (228, 304, 4), (1080, 1920, 24), (24, 512, 70),
                                   (2000, 2000, 100), (3, 4, 2), (700, 64, 4), (1, 4, 1), (641, 128, 56), (40, 4096, 3000)])
def test_plan_geometry_for_named_shapes(H, W, n):
    check_plan(H, W, n)


@settings(max_examples=150, deadline=None)
@given(H=st.integers(1, 3000), W4=st.integers(1, 700), n=st.integers(1, 400))
def test_plan_geometry_for_random_shapes(H, W4, n):
    check_plan(H, 4 * W4, n)
    check_plan(H, 4 * W4, n, chained=True)


def test_headline_shape_chains_into_twelve_strips():
    info = check_plan(352, 1216, 24, chained=True)
    (p,) = info['passes']
    assert p['count'] == 1 and len(p['strips']) == 12 and p['strips'][1][0] == 104


def test_headline_shape_is_a_single_pass_single_band_plan():
    info = check_plan(352, 1216, 24)
    (p,) = info['passes']
    assert p['count'] == 1 and len(p['bands']) == 1 and p['cs'] * p['RB'] >= 352 and len(p['strips']) == 15


def replay_chained(info, guidance, blur, sparse, norm):
    """Chained plan on the numpy oracle: strips left to right; a strip's column -1 is, at every step, what the strip to its
    left computed there (its history), so only the right edge goes stale."""
    f64 = np.float64
    gate_wb, gate_sum = onp.affinity_normalization_2d(guidance.astype(f64), norm)
    raw = blur.astype(f64)
    mask = None if sparse is None else np.sign(sparse.astype(f64))
    H, W = raw.shape[-2:]
    cur = raw
    for p in info['passes']:
        for _ in range(p['count']):
            nxt = np.full_like(cur, np.nan)
            hist = None                                          # [step] -> column (..., H, 1) left of the current strip
            for i, (tx0, ux0, ux1) in enumerate(p['strips']):
                xs = slice(tx0, min(tx0 + p['TW'], W))
                d, d0 = cur[..., xs], raw[..., xs]
                w, s = gate_wb[..., xs], gate_sum[..., xs]
                m = None if mask is None else mask[..., xs]
                rec = ux1 - 1 - tx0                              # tile column recorded for the next strip
                new_hist = []
                with np.errstate(invalid='ignore'):
                    for t in range(p['iters']):
                        new_hist.append(d[..., rec:rec + 1].copy())
                        left = hist[t] if hist is not None else np.zeros_like(d[..., :1])
                        ext = np.concatenate([left, d], axis=-1)             # column -1 in front: exact, not zero
                        acc = np.zeros_like(d)
                        for k, (dy, dx) in enumerate(onp.OFFSETS_2D):
                            acc = acc + w[:, k:k + 1] * onp.shift2d(ext, dy, dx)[..., 1:]
                        d = (1.0 - s) * d0 + acc
                        if m is not None:
                            d = (1.0 - m) * d + m * d0
                nxt[..., ux0:ux1] = d[..., ux0 - tx0:ux1 - tx0]
                hist = new_hist
            cur = nxt
    return cur


def replay(info, guidance, blur, sparse, norm):
    """Runs the plan on the numpy oracle in float64, tile by tile."""
    f64 = np.float64
    gate_wb, gate_sum = onp.affinity_normalization_2d(guidance.astype(f64), norm)
    raw = blur.astype(f64)
    mask = None if sparse is None else np.sign(sparse.astype(f64))
    H, W = raw.shape[-2:]
    cur = raw
    for p in info['passes']:
        CH = p['cs'] * p['RB']
        for _ in range(p['count']):
            nxt = np.full_like(cur, np.nan)
            for by0, uy0, uy1 in p['bands']:
                for tx0, ux0, ux1 in p['strips']:
                    ys, xs = slice(by0, min(by0 + CH, H)), slice(tx0, min(tx0 + p['TW'], W))
                    d, d0 = cur[..., ys, xs], raw[..., ys, xs]
                    w, s = gate_wb[..., ys, xs], gate_sum[..., ys, xs]
                    m = None if mask is None else mask[..., ys, xs]
                    with np.errstate(invalid='ignore'):
                        for _ in range(p['iters']):            # zero outside the tile: shift2d pads with zeros
                            acc = np.zeros_like(d)
                            for k, (dy, dx) in enumerate(onp.OFFSETS_2D):
                                acc = acc + w[:, k:k + 1] * onp.shift2d(d, dy, dx)
                            d = (1.0 - s) * d0 + acc
                            if m is not None:
                                d = (1.0 - m) * d + m * d0
                    nxt[..., uy0:uy1, ux0:ux1] = d[..., uy0 - by0:uy1 - by0, ux0 - tx0:ux1 - tx0]
            cur = nxt
    return cur


@pytest.mark.parametrize('H,W,n,norm,with_sparse', [
    (50, 300, 10, '8sum', True),        # strips with halos, one band
    (60, 260, 45, '8sum_abs', False),   # several passes
    (700, 64, 6, '8sum', True),         # row bands
    (660, 260, 20, '8sum', True),       # bands x strips
    (33, 132, 131, '8sum', True),       # passes of unequal length
])
def test_replaying_the_plan_on_the_oracle_reproduces_the_untiled_result(H, W, n, norm, with_sparse):
    rng = np.random.default_rng(H * 31 + W + n)
    g = rng.standard_normal((1, 8, H, W))
    d = rng.uniform(0, 10, (1, 2, H, W))
    s = None
    if with_sparse:
        s = rng.uniform(-1, 10, (1, 1, H, W)) * (rng.uniform(size=(1, 1, H, W)) < 0.02)
    info = check_plan(H, W, n)
    ref = onp.cspn2d(g, d, s, n, norm, dtype=np.float64)
    out = replay(info, g, d, s, norm)
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    cinfo = check_plan(H, W, n, chained=True)
    if cinfo['supported']:
        out = replay_chained(cinfo, g, d, s, norm)
        assert not np.isnan(out).any()
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
