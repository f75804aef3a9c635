"""Pure-Python restatements of two pieces of index arithmetic the CUDA kernels rely on, checked exhaustively on the CPU
(the GPU parity tests cover the kernels themselves; these pin the reasoning their comments give)."""
import numpy as np
import pytest


def _shfl_xor(v, m):
    return v[np.arange(32) ^ m]


@pytest.mark.parametrize("W", [4, 8, 16])
def test_select_free_reduce_scatter_slot_order(W):
    """composite_bwd.cu: reduce_scatter_permuted.  Lane l keeps component c in slot  c ^ (~(l >> (5 - log2 W)) & (W - 1));
    every level sends slots [0, n/2) and keeps [n/2, n) with NO per-lane select, and lane l must end up with the
    32-lane total of component  l >> (5 - log2 W)."""
    rng = np.random.default_rng(W)
    vals = rng.standard_normal((32, W))                       # vals[lane, component]
    lg = W.bit_length() - 1
    lanes = np.arange(32)
    K = ~(lanes >> (5 - lg)) & (W - 1)
    x = np.empty((W, 32))                                      # x[slot] is a 32-lane register
    for p in range(W):
        x[p] = vals[lanes, p ^ K]
    m, n = 16, W
    while n > 1:
        for i in range(n // 2):
            x[i] = x[i + n // 2] + _shfl_xor(x[i], m)
        m >>= 1
        n >>= 1
    r = x[0]
    mm = 16 // W
    while mm > 0:
        r = r + _shfl_xor(r, mm)
        mm >>= 1
    want = vals.sum(axis=0)[lanes >> (5 - lg)]
    assert np.allclose(r, want, rtol=1e-12, atol=1e-12)


def test_gradient_row_lane_assignment_is_collision_free():
    """composite_bwd.cu: which lane adds which finished component to the Gaussian's gradient row (8 geometry slots + 4*NG
    channels in chunks of 16 / 8 / 4).  No lane carries two components, every real slot is carried exactly once."""
    for NG in range(1, 8):
        NC = 4 * NG
        C16 = 16 if NC >= 16 else 0
        C8 = 8 if NC - C16 >= 8 else 0
        C4 = NC - C16 - C8
        B8, B4 = C16, C16 + C8
        second = C16 > 0 and C8 > 0 and C4 > 0
        owner = {}
        for lane in range(32):
            sel, off = -1, 0
            if lane & 3 == 0:
                if lane >> 2 != 7:
                    sel, off = 0, lane >> 2
            elif C16 and lane & 1:
                sel, off = 1, 8 + (lane >> 1)
            elif C8 and lane & 3 == 2:
                sel, off = 2, 8 + B8 + (lane >> 2)
            if C4 and not second and lane & 7 == (2 if C16 else 1):
                assert sel == -1, (NG, lane)                    # the 4-chunk's lane class must be free
                sel, off = 3, 8 + B4 + (lane >> 3)
            if sel >= 0:
                assert off not in owner, (NG, lane, off)
                owner[off] = lane
        if second:                                              # NG == 7: the 4-chunk goes out with a second instruction
            for lane in range(0, 32, 8):
                owner[8 + B4 + (lane >> 3)] = lane
        assert sorted(owner) == list(range(7)) + list(range(8, 8 + NC)), NG


def test_block_mask_word_split_covers_each_instance_once():
    """composite.cu: block_mask_kernel writes a tile's [x, y) instances as <= 3 head bytes, whole 4-instance words and
    <= 3 tail bytes; neighbouring tiles share the boundary words, so no byte may be written by two tiles or twice."""
    for x in range(0, 13):
        for y in range(x, x + 14):
            a0 = min((x + 3) & ~3, y)
            a1 = max(y & ~3, a0)
            written = []
            for t in range(8):
                i = x + t if t < 4 else a1 + (t - 4)
                mine = i < a0 if t < 4 else i < y
                if mine:
                    written.append(i)
            for w in range(a0 >> 2, a1 >> 2):
                written += [4 * w, 4 * w + 1, 4 * w + 2, 4 * w + 3]
            assert sorted(written) == list(range(x, y)), (x, y, written)
