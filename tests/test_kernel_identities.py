"""Arithmetic identities the CUDA kernels rely on, checked exhaustively / on tie-heavy random inputs on the CPU.

These are not parity tests (the `-m gpu` tests compare the kernels with the oracle); they pin the REWRITES that keep the
kernels' instruction counts down, so that a later edit of either side shows up without a GPU:

* pool backward (`pool22_bwd_strip_kernel`, kernels_fast.cu): "element e is the FIRST maximum of a 2x2 window in row-major
  order (a later member displaces an earlier one only when strictly greater — TF's MaxPoolGrad tie rule)"
  is decided by compares against maxima of the other members instead of a 4-way arg-max;
* head backward (`dense_head_bwd_kernel`): `i / N` as `(i * ceil(65536 / N)) >> 16` for the CTA's slab of W;
* strip kernels: the thread index is split with 32-bit unsigned div/mod.
"""
import numpy as np
import pytest


def first_max_position(p0, p1, p2, p3):
    """the reference rule, as the kernels' former arg-max: strict > displaces the earlier member"""
    k, best = 0, p0
    if p1 > best:
        k, best = 1, p1
    if p2 > best:
        k, best = 2, p2
    if p3 > best:
        k = 3
    return k


def wins(pos, p):
    """the rewrite: beat the earlier members strictly and the later ones weakly, through maxima of the others"""
    e = p[pos]
    if pos == 0:
        return e >= max(p[1], p[2], p[3])
    if pos == 1:
        return e > p[0] and e >= max(p[2], p[3])
    if pos == 2:
        return e > max(p[0], p[1]) and e >= p[3]
    return e > max(p[0], p[1], p[2])


def test_first_maximum_by_neighbour_maxima_equals_the_argmax_rule_exhaustively_on_ties():
    # values from a 4-element set: every tie pattern of a window occurs (4^4 windows), plus zeros as ReLU produces them
    vals = [0.0, 0.5, 0.5000001, 2.0]
    for a in vals:
        for b in vals:
            for c in vals:
                for d in vals:
                    p = (a, b, c, d)
                    k = first_max_position(*p)
                    got = [wins(pos, p) for pos in range(4)]
                    assert got == [pos == k for pos in range(4)], (p, k, got)


def test_pool_backward_walk_restated_with_the_rewrite_matches_the_oracle():
    """the kernel's per-element formulation (4 windows around (h, w), fixed summation order) against the oracle's pool gradient
    on a tie-heavy activation map (post-ReLU zeros)"""
    rng = np.random.Generator(np.random.PCG64(5))
    x = np.maximum(rng.integers(-3, 4, size=(2, 6, 7, 3)).astype(np.float32), 0.0)  # many exact ties
    dy = rng.standard_normal((2, 5, 6, 3)).astype(np.float32)
    from oracle.layers import pool2d_grad  # MaxPoolGradV2 restated (golden: nn/KernelsSpec.scala:277-288)

    ref = pool2d_grad(x, dy, (2, 2), (1, 1), "valid")
    N, H, W, C = x.shape
    got = np.zeros_like(x)
    for n in range(N):
        for h in range(H):
            for w in range(W):
                for c in range(C):
                    acc = np.float32(0)
                    # windows (h-1,w-1), (h-1,w), (h,w-1), (h,w): the element is member 3, 2, 1, 0 of them
                    for (wh, ww, pos) in ((h - 1, w - 1, 3), (h - 1, w, 2), (h, w - 1, 1), (h, w, 0)):
                        if 0 <= wh < H - 1 and 0 <= ww < W - 1:
                            p = (x[n, wh, ww, c], x[n, wh, ww + 1, c], x[n, wh + 1, ww, c], x[n, wh + 1, ww + 1, c])
                            if wins(pos, p):
                                acc = np.float32(acc + dy[n, wh, ww, c])
                    got[n, h, w, c] = acc
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("n", range(1, 17))
def test_multiply_shift_row_split_is_exact_for_the_head_slab(n):
    # dense_head_bwd_kernel: slab of at most 64 k columns x N <= 16 classes
    magic = (65536 + n - 1) // n
    i = np.arange(64 * 16, dtype=np.uint64)
    np.testing.assert_array_equal((i * magic) >> 16, i // n)


def test_32_bit_thread_index_split_matches_the_64_bit_one():
    # (i % C4, (i / C4) % segs, (i / C4 / segs) % rows, i / C4 / segs / rows) with i < 2^31 (the launchers' bound)
    rng = np.random.Generator(np.random.PCG64(6))
    for _ in range(200):
        c4, segs, rows = int(rng.integers(1, 65)), int(rng.integers(1, 9)), int(rng.integers(1, 300))
        i = rng.integers(0, 2**31 - 1, size=64, dtype=np.int64)
        u = i.astype(np.uint32)
        t64, t32 = i // c4, u // np.uint32(c4)
        assert np.array_equal(i % c4, (u % np.uint32(c4)).astype(np.int64))
        assert np.array_equal(t64 % segs, (t32 % np.uint32(segs)).astype(np.int64))
        assert np.array_equal((t64 // segs) % rows, ((t32 // np.uint32(segs)) % np.uint32(rows)).astype(np.int64))
        assert np.array_equal(t64 // segs // rows, (t32 // np.uint32(segs) // np.uint32(rows)).astype(np.int64))


def test_vector_row_loads_of_the_fused_conv_cover_exactly_the_in_row_columns():
    # conv_smallk_pool_fwd_kernel, single-channel rows with W % 4 == 0, no left padding: a thread of segment ow0 = 4*seg needs the
    # PW + KW - 1 = 7 inputs ow0 .. ow0+6 (zero beyond the row); it loads quad ow0..ow0+3 iff ow0 + 4 <= W and quad ow0+4..ow0+7
    # iff ow0 + 8 <= W.  The quads are never split by the row end, so this equals the bounds-checked scalar loads -- also with
    # explicit right padding, where output columns (and whole segments) lie beyond the input row.
    for W in range(4, 132, 4):
        for pad_right in (0, 1, 3, 6):
            OW = W + pad_right - 2  # 3x3, stride 1
            for seg in range((OW + 3) // 4):
                ow0 = 4 * seg
                scalar = [(ow0 + q) if ow0 + q < W else None for q in range(7)]
                lo = [ow0 + j for j in range(4)] if ow0 + 4 <= W else [None] * 4
                hi = [ow0 + 4 + j for j in range(4)] if ow0 + 8 <= W else [None] * 4
                assert (lo + hi)[:7] == scalar, (W, pad_right, seg)


def win4(p, relu, thr):
    """kernels_fast.cu `win4`: winner byte of a window = index of the first maximum | (passes the ReLU) << 2"""
    k = first_max_position(*p)
    return k | (4 if (not relu or p[k] > thr) else 0)


def vcmpeq4(a, b):
    out = 0
    for s in (0, 8, 16, 24):
        if (a >> s) & 0xFF == (b >> s) & 0xFF:
            out |= 0xFF << s
    return out


def take4(m, idx, relu):
    """kernels_fast.cu `take4`: per-byte routing test on a packed word of four channels' winner bytes"""
    return vcmpeq4(m, 0x01010101 * (idx | 4)) if relu else vcmpeq4(m & 0x03030303, 0x01010101 * idx)


@pytest.mark.parametrize("relu", [False, True])
def test_winner_map_routing_matches_the_oracle_pool_and_relu_gradient(relu):
    """pool22_bwd_map_kernel restated: the forward records win4 per window, the backward routes dy with take4 on packed words
    (0xffffffff = no window) in the order (h-1,w-1), (h-1,w), (h,w-1), (h,w) -- against the oracle's MaxPoolGrad followed by
    the ReLU gradient, on a tie-heavy post-ReLU map"""
    from oracle.layers import pool2d_grad
    rng = np.random.Generator(np.random.PCG64(11))
    thr = 0.0
    z = rng.integers(-3, 4, size=(2, 6, 7, 4)).astype(np.float32)
    x = np.maximum(z, thr) if relu else z
    dy = rng.standard_normal((2, 5, 6, 4)).astype(np.float32)
    ref = pool2d_grad(x, dy, (2, 2), (1, 1), "valid")
    if relu:
        ref = ref * (x > thr)
    N, H, W, C = x.shape
    words = np.zeros((N, H - 1, W - 1), dtype=np.uint64)
    for n in range(N):
        for h in range(H - 1):
            for w in range(W - 1):
                word = 0
                for c in range(4):
                    p = (x[n, h, w, c], x[n, h, w + 1, c], x[n, h + 1, w, c], x[n, h + 1, w + 1, c])
                    word |= win4(p, relu, thr) << (8 * c)
                words[n, h, w] = word
    got = np.zeros_like(x)
    for n in range(N):
        for h in range(H):
            for w in range(W):
                acc = [np.float32(0)] * 4
                for (wh, ww, pos) in ((h - 1, w - 1, 3), (h - 1, w, 2), (h, w - 1, 1), (h, w, 0)):
                    ok = 0 <= wh < H - 1 and 0 <= ww < W - 1
                    m = int(words[n, wh, ww]) if ok else 0xFFFFFFFF
                    g = dy[n, wh, ww] if ok else np.zeros(4, np.float32)
                    e = take4(m, pos, relu)
                    for c in range(4):
                        if e & (0xFF << (8 * c)):
                            acc[c] = np.float32(acc[c] + g[c])
                got[n, h, w] = acc
    np.testing.assert_array_equal(got, ref)
