"""Property tests of the oracle (hypothesis): the invariants the GPU tests rely on at full size are true of the
arithmetic itself, for arbitrary small shapes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle, cspn_numpy as onp

shapes = st.tuples(st.integers(1, 2), st.integers(1, 2), st.integers(1, 9), st.integers(2, 11))   # W >= 2: every pixel has a neighbour
# (a 1x1 image has no neighbour at all: 0/0 affinities, NaN -- pinned separately by tests/golden/tiny_1x1.npz)


def inputs(seed, B, C, H, W, sparse_frac=0.2):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((B, 8, H, W)).astype(np.float32)
    d = (rng.random((B, C, H, W)) * 10).astype(np.float32)
    s = ((rng.random((B, 1, H, W)) < sparse_frac) * rng.standard_normal((B, 1, H, W))).astype(np.float32)
    return g, d, s


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(0, 6), st.sampled_from(['8sum', '8sum_abs']), st.integers(0, 10_000))
def test_c_and_numpy_oracles_agree(shape, n, norm, seed):
    g, d, s = inputs(seed, *shape)
    a = onp.cspn2d(g, d, s, n, norm)
    b = c_oracle.cspn2d(g, d, s, n, norm)
    assert onp.parity_ok(b, a, 1e-5)[0]


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(1, 6), st.sampled_from(['8sum', '8sum_abs']), st.integers(0, 10_000))
def test_output_is_linear_in_blur_depth(shape, n, norm, seed):
    g, d, s = inputs(seed, *shape)
    d = d.astype(np.float64)
    d2 = np.random.default_rng(seed + 1).random(d.shape)
    f = lambda x: onp.cspn2d(g, x, s, n, norm, dtype=np.float64)
    lhs = f(0.5 * d - 3 * d2)
    rhs = 0.5 * f(d) - 3 * f(d2)
    assert np.allclose(lhs, rhs, rtol=1e-8, atol=1e-8, equal_nan=True)


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(1, 6), st.integers(0, 10_000))
def test_abs_mode_preserves_constants_and_batch_items_are_independent(shape, n, seed):
    g, d, s = inputs(seed, *shape)
    const = np.full_like(d, 2.5)
    out = onp.cspn2d(g, const, s, n, '8sum_abs', dtype=np.float64)
    assert np.allclose(out, 2.5, rtol=1e-9)
    full = onp.cspn2d(g, d, s, n, '8sum')
    one = onp.cspn2d(g[:1], d[:1], s[:1], n, '8sum')
    assert np.array_equal(full[:1], one)


@settings(max_examples=25, deadline=None)
@given(shapes, st.integers(1, 6), st.sampled_from(['8sum', '8sum_abs']), st.integers(0, 10_000))
def test_positive_sparse_pixels_return_the_blur_depth(shape, n, norm, seed):
    g, d, s = inputs(seed, *shape, sparse_frac=0.5)
    out = onp.cspn2d(g, d, s, n, norm)
    m = (s > 0) & np.isfinite(out)
    assert np.array_equal(np.broadcast_to(m, out.shape) * out, np.broadcast_to(m, out.shape) * d)
