"""Index arithmetic of the dense-x conv kernel (csrc/az_net_tc.cu k_conv_tc_x, AZ_TC_MODE=5) restated in numpy: the 3-D TMA
boxes with out-of-bounds zero fill, the row-shifted operand views and the tile / validity bookkeeping must add up to an
ordinary zero-padded 3x3 convolution.  This pins the geometry (the part that cannot be seen in a profiler); the tensor-core
plumbing itself is checked on the GPU by tests/tc_mode_check.py."""
import numpy as np
import pytest


def _tma_box(act3, c0, x0, br0, nbr, hc):
    """cp.async.bulk.tensor.3d of box {64, hc, nbr} at (c0, x0, br0) from act3[board_row][x][channel], zeros out of bounds."""
    nb_rows, width, _ = act3.shape
    box = np.zeros((nbr, hc, 64), dtype=act3.dtype)
    for j in range(nbr):
        for i in range(hc):
            b, x = br0 + j, x0 + i
            if 0 <= b < nb_rows and 0 <= x < width:
                box[j, i] = act3[b, x, c0:c0 + 64]
    return box.reshape(nbr * hc, 64)  # shared-memory rows, 128 bytes each


@pytest.mark.parametrize('hc,leaves,cin', [(9, 7, 64), (9, 3, 128), (17, 2, 64), (5, 11, 64), (13, 4, 64), (19, 3, 64)])
def test_dense_x_equals_zero_padded_conv(hc, leaves, cin):
    rng = np.random.default_rng(hc * 100 + leaves)
    cout = 8
    RP, guard = hc * (hc + 1), hc * 16
    TH = 256 // hc * hc
    nbr = (256 + 2 * hc + hc - 1) // hc
    rows_total = leaves * (hc + 1) * (hc + 1) + 2 * guard + 1024  # the engine sizes the buffer for the wider halo layout
    act = np.zeros((rows_total, cin), dtype=np.float64)
    x = rng.standard_normal((leaves, hc, hc, cin))
    for lf in range(leaves):
        for y in range(hc):
            act[guard + lf * RP + y * hc: guard + lf * RP + y * hc + hc] = x[lf, y]
    # rows past the live leaves hold stale values of an earlier, larger batch: they must not reach any valid output
    act[guard + leaves * RP:] = rng.standard_normal((rows_total - guard - leaves * RP, cin))
    act[guard + leaves * RP: guard + leaves * RP + 0] = 0
    w = rng.standard_normal((9, cout, cin))
    act3 = act[: rows_total // hc * hc].reshape(rows_total // hc, hc, cin)
    M = leaves * RP
    out = np.zeros((rows_total, cout))
    for t in range((M + TH - 1) // TH):
        br0 = (guard + t * TH) // hc - 1
        acc = np.zeros((256, cout))
        for kk in range(cin // 64):
            for dxi in range(3):
                sub = _tma_box(act3, kk * 64, dxi - 1, br0, nbr, hc)
                for dyi in range(3):
                    tap = dyi * 3 + dxi
                    row0 = hc + (dyi - 1) * hc  # descriptor start: output row 0 sits one board row into the box
                    assert row0 >= 0 and row0 + 256 <= sub.shape[0]
                    acc += sub[row0:row0 + 256] @ w[tap][:, kk * 64:(kk + 1) * 64].T
        for l in range(256):
            m = t * TH + l
            if l < TH and m < M:
                valid = (m % RP) // hc < hc
                out[guard + m] = acc[l] if valid else 0.0
    # reference: zero-padded 3x3 correlation per leaf (torch conv2d semantics, weight[co][ci][ky][kx] <-> w[ky*3+kx][co][ci])
    for lf in range(leaves):
        xp = np.zeros((hc + 2, hc + 2, cin))
        xp[1:-1, 1:-1] = x[lf]
        for y in range(hc):
            for xx in range(hc):
                ref = sum(xp[y + ky, xx + kx] @ w[ky * 3 + kx].T for ky in range(3) for kx in range(3))
                np.testing.assert_allclose(out[guard + lf * RP + y * hc + xx], ref, atol=1e-9)
        assert not out[guard + lf * RP + hc * hc: guard + (lf + 1) * RP].any()  # the leaf's zero board row stays zero
