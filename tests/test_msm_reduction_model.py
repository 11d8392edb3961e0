"""CPU: the index arithmetic of the MSM bucket reduction (csrc/msm.cu: k_msm_rowcol, k_msm_rc_bits, k_msm_final) and of the
window-width model, restated on integers.  Group addition is linear, so checking the weights on integers (every bucket
holding a random integer instead of a point) checks the decomposition the kernels implement."""
import random

import pytest


def rowcol_items(buckets, c):
    """items[e] (weight 2^e) exactly as the kernels produce them: row / column sums, then bit-slicing"""
    nb = len(buckets)
    assert nb == 1 << (c - 1)
    lb = c // 2                                   # (c - 1 + 1) / 2 in run_msm
    nC, nR = 1 << lb, nb >> lb
    R = [sum(buckets[hi * nC + lo] for lo in range(nC)) for hi in range(nR)]
    C = [sum(buckets[hi * nC + lo] for hi in range(nR)) for lo in range(nC)]
    items = []
    for e in range(c - 1):
        v = 0
        if e <= lb:
            v += sum(C[lo] for lo in range(nC) if ((lo + 1) >> e) & 1)
        if e >= lb:
            v += sum(R[hi] for hi in range(nR) if (hi >> (e - lb)) & 1)
        items.append(v)
    return items


@pytest.mark.parametrize("c", [9, 10, 11, 13, 15, 16])
def test_row_column_reduction_weights(c):
    rnd = random.Random(c)
    nb = 1 << (c - 1)
    buckets = [rnd.randrange(1 << 64) for _ in range(nb)]
    want = sum((b + 1) * v for b, v in enumerate(buckets))          # bucket b holds the points with |digit| = b + 1
    items = rowcol_items(buckets, c)
    assert len(items) == c - 1
    assert sum(v << e for e, v in enumerate(items)) == want
    # a single non-empty bucket, first and last
    for b in (0, nb - 1, nb // 2 + 3):
        one = [0] * nb
        one[b] = 1
        assert sum(v << e for e, v in enumerate(rowcol_items(one, c))) == b + 1


def table_window_bits(k):
    """csrc/msm.cu table_window_bits"""
    c = max(4, min(16, k - 3))
    if k < 14:
        return c
    work = lambda cc: (254 // cc + 1) * (1 << k) * 10.0 + (1 << (cc - 1)) * 5.1 * 14.0
    best = min(work(cc) for cc in range(c, 17))
    while c < 16 and work(c) > 1.03 * best:
        c += 1
    return c


def test_window_width_model_matches_the_measured_choices():
    # measured best c on B200 at k = 16, 17, 18 (complete proofs), cap 16 from k = 19
    assert [table_window_bits(k) for k in (16, 17, 18, 19, 20, 24)] == [15, 15, 16, 16, 16, 16]
    for k in range(4, 28):
        c = table_window_bits(k)
        assert 4 <= c <= 16
        # the shared-memory digit sort needs 2^(c-1) counters of 4 bytes in <= 200 KB
        assert (1 << (c - 1)) * 4 <= 200 * 1024


def test_signed_digits_cover_the_scalar():
    """W = floor(254 / c) + 1 signed digits in (-2^(c-1), 2^(c-1)] reproduce every 254-bit scalar (no carry out of the top
    window); bucket index = |digit| - 1 < 2^(c-1)"""
    rnd = random.Random(7)
    r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
    for c in (13, 15, 16):
        W = 254 // c + 1
        for s in [0, 1, r - 1, (1 << 253) + 12345] + [rnd.randrange(r) for _ in range(200)]:
            digits, carry, x = [], 0, s
            for w in range(W):
                d = ((x >> (c * w)) & ((1 << c) - 1)) + carry
                carry = 0
                if d > (1 << (c - 1)):
                    d -= 1 << c
                    carry = 1
                digits.append(d)
            assert carry == 0
            assert sum(d << (c * w) for w, d in enumerate(digits)) == s
            assert all(-(1 << (c - 1)) < d <= (1 << (c - 1)) for d in digits)
