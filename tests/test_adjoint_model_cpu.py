"""A numpy model of one adjoint step of cspn2d_cluster.cu (kAdjoint mode), written line for line after
adj_row / fold_edges / iterate_adj and the exchange-slot convention of publish(): thread = PR x PC patch, lanes along x,
warps along y, partial rows for the neighbouring warps through slots 1+2wy / 2+2wy, read back from 2wy / 2wy+3.
It must equal the transpose of the forward step,  lambda_t(q) = sum_k w_k(q - off_k) lambda_{t+1}(q - off_k)."""
import numpy as np
import pytest

from oracle.cspn_numpy import OFFSETS_2D, shift2d


def tap_of(dy, dx):
    return 1 - dx if dy == 1 else ((3 if dx == 1 else 4) if dy == 0 else 6 - dx)


def test_tap_of_is_the_channel_of_the_offset():
    for k, off in enumerate(OFFSETS_2D):
        assert tap_of(*off) == k


def adj_row(w, l, acc, DY):
    """w[PC][8], l[PC] of one SOURCE row; pushes into destination row DY further down; returns (xl, xr)."""
    PC = len(l)
    xl = xr = 0.0
    for j in range(PC):
        for dx in (1, 0, -1):
            if DY == 0 and dx == 0:
                continue
            jd = j + dx
            c = w[j][tap_of(DY, dx)] * l[j]
            if jd < 0:
                xl += c
            elif jd >= PC:
                xr += c
            else:
                acc[jd] += c
    return xl, xr


def fold_edges(rows, xl, xr):
    """rows[lane][PC]; xl/xr[lane]: lane's contribution to lane-1's last column / lane+1's first column."""
    L = len(rows)
    for lane in range(L):
        if lane > 0:
            rows[lane][0] += xr[lane - 1]           # __shfl_up(xr, 1)
        if lane < L - 1:
            rows[lane][-1] += xl[lane + 1]          # __shfl_down(xl, 1)


def model_step(w, lam, NW, PR, L, PC):
    """w[H][W][8], lam[H][W] with H = NW*PR, W = L*PC -> lambda of the previous step, via the kernel's decomposition."""
    slots = np.zeros((2 * NW + 2, L * PC))
    out = np.zeros_like(lam)
    patch = lambda a, wy, lane, r: a[wy * PR + r, lane * PC:(lane + 1) * PC]
    # phase 1: every warp publishes the partial rows it owes to the rows above / below its patch
    for wy in range(NW):
        up = [np.zeros(PC) for _ in range(L)]
        dn = [np.zeros(PC) for _ in range(L)]
        upl, upr, dnl, dnr = np.zeros(L), np.zeros(L), np.zeros(L), np.zeros(L)
        for lane in range(L):
            upl[lane], upr[lane] = adj_row(patch(w, wy, lane, 0), patch(lam, wy, lane, 0), up[lane], -1)
            dnl[lane], dnr[lane] = adj_row(patch(w, wy, lane, PR - 1), patch(lam, wy, lane, PR - 1), dn[lane], +1)
        fold_edges(up, upl, upr)
        fold_edges(dn, dnl, dnr)
        slots[1 + 2 * wy] = np.concatenate(up)
        slots[2 + 2 * wy] = np.concatenate(dn)
    # phase 2: own contributions, x-edges, then the neighbours' rows
    for wy in range(NW):
        lout = [[np.zeros(PC) for _ in range(L)] for _ in range(PR)]
        xl, xr = np.zeros((PR, L)), np.zeros((PR, L))
        for lane in range(L):
            for r in range(PR):
                wr, lr = patch(w, wy, lane, r), patch(lam, wy, lane, r)
                for DY in (-1, 0, 1):
                    if 0 <= r + DY < PR:
                        a, b = adj_row(wr, lr, lout[r + DY][lane], DY)
                        xl[r + DY, lane] += a
                        xr[r + DY, lane] += b
        for r in range(PR):
            fold_edges(lout[r], xl[r], xr[r])
        rows = [np.concatenate(lout[r]) for r in range(PR)]
        rows[0] = rows[0] + slots[2 * wy]             # what the warp above owes my top row (zeros at the CTA edge)
        rows[PR - 1] = rows[PR - 1] + slots[2 * wy + 3]
        for r in range(PR):
            out[wy * PR + r] = rows[r]
    return out


@pytest.mark.parametrize('NW,PR,L,PC', [(2, 2, 3, 4), (3, 5, 4, 4), (1, 2, 2, 4), (4, 3, 2, 4)])
def test_model_of_the_adjoint_step_is_the_transpose_of_the_forward_step(NW, PR, L, PC):
    rng = np.random.default_rng(NW * 100 + PR * 10 + L)
    H, W = NW * PR, L * PC
    w = rng.standard_normal((H, W, 8))
    lam = rng.standard_normal((H, W))
    got = model_step(w, lam, NW, PR, L, PC)
    # definition: lambda_t(q) = sum_k (w_k lambda_{t+1})(q - off_k), zero outside the tile
    want = sum(shift2d(w[:, :, k] * lam, -dy, -dx) for k, (dy, dx) in enumerate(OFFSETS_2D))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    # and it is the transpose of the forward stencil: <A d, lam> == <d, A^T lam>
    d = rng.standard_normal((H, W))
    fwd = sum(w[:, :, k] * shift2d(d, dy, dx) for k, (dy, dx) in enumerate(OFFSETS_2D))
    assert np.isclose((fwd * lam).sum(), (d * got).sum(), rtol=1e-10)
