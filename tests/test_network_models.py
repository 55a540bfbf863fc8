"""CPU models of the index maps behind two hand-written device networks, kept next to the kernels they describe:
 * tile_sort.cu -- the register-blocked all-ascending bitonic network (64-bit words) and its pair-word 32-bit variant:
   the models run the same group decomposition (levels 2/4/8[/16] in registers, flip group, j groups of <= 3 stages,
   virtual +inf padding) and must sort every length; the skewed shared-memory layout must make a thread's eight
   addresses `base + constant` (that is what the immediates in the kernel assume);
 * blend_backward.cu -- the 10-value transposing warp reduction and the lane -> value map `red10_index`.
They do not exercise the GPU (the GPU parity tests do); they pin the arithmetic the kernels were derived from."""
import random

INF = float("inf")


def _ce(v, x, y):
    if v[x] > v[y]:
        v[x], v[y] = v[y], v[x]


def _ld(a, i, n):
    return a[i] if i < n else INF


def _st(a, i, n, v):
    if i < n:
        a[i] = v
    else:
        assert v == INF                 # a comparator never moves a real word beyond n


def _ilog2(x):
    return x.bit_length() - 1


def _sort64_model(a, cap):
    n = len(a)
    g = 0
    while g * 8 < n:                    # levels 2, 4, 8 on 8 consecutive words
        v = [_ld(a, 8 * g + m, n) for m in range(8)]
        for p in [(0, 1), (2, 3), (4, 5), (6, 7), (0, 3), (1, 2), (4, 7), (5, 6), (0, 1), (2, 3), (4, 5), (6, 7),
                  (0, 7), (1, 6), (2, 5), (3, 4), (0, 2), (1, 3), (4, 6), (5, 7), (0, 1), (2, 3), (4, 5), (6, 7)]:
            _ce(v, *p)
        for m in range(8):
            _st(a, 8 * g + m, n, v[m])
        g += 1

    def flip_group(K):
        S = K // 8
        for g in range(((n + K - 1) // K) * S):
            r, base = g & (S - 1), (g // S) * K
            il, iu = base + r, base + K - 1 - r
            if il >= n:
                continue
            v = [_ld(a, il + S * m, n) for m in range(4)] + [_ld(a, iu - S * m, n) for m in range(4)]
            for m in range(4):
                _ce(v, m, 4 + m)
            for p in [(0, 2), (1, 3), (6, 4), (7, 5), (0, 1), (2, 3), (5, 4), (7, 6)]:
                _ce(v, *p)
            for m in range(4):
                _st(a, il + S * m, n, v[m]); _st(a, iu - S * m, n, v[4 + m])

    def j_stages_from(J):
        while J >= 1:
            cnt = min(3, _ilog2(J) + 1)
            S = J >> (cnt - 1)
            if S < n:
                for g in range(((n + 8 * S - 1) // (8 * S)) * S):
                    i0 = (g // S) * 8 * S + (g & (S - 1))
                    if i0 + S >= n:
                        continue
                    v = [_ld(a, i0 + S * m, n) for m in range(8)]
                    if cnt >= 3:
                        for p in [(0, 4), (1, 5), (2, 6), (3, 7)]: _ce(v, *p)
                    if cnt >= 2:
                        for p in [(0, 2), (1, 3), (4, 6), (5, 7)]: _ce(v, *p)
                    for p in [(0, 1), (2, 3), (4, 5), (6, 7)]: _ce(v, *p)
                    for m in range(8):
                        _st(a, i0 + S * m, n, v[m])
            J = S // 2

    K = 16
    while K <= cap and K // 2 < n:
        flip_group(K); j_stages_from(K // 16); K *= 2
    return a


def test_register_blocked_bitonic_network_sorts_every_length():
    rng = random.Random(1)
    for n in list(range(1, 70)) + [100, 127, 128, 129, 255, 257, 500, 1000, 1023, 1025, 1100, 2047]:
        a = [rng.randrange(0, 50) for _ in range(n)]        # many ties
        assert _sort64_model(list(a), 2048) == sorted(a), n


def test_skewed_layout_gives_constant_offsets():
    for shift in (4,):                                      # 64-bit words: one pad word per 16
        phys = lambda i: i + (i >> shift)
        for S in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048]:
            for g in range(2048):
                i0 = (g // S) * 8 * S + (g & (S - 1))
                assert all(phys(i0 + S * m) == phys(i0) + phys(S * m) for m in range(8))
        for K in [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]:
            S = K // 8
            for g in range(2048):
                r, base = g & (S - 1), (g // S) * K
                il, iu = base + r, base + K - 1 - r
                assert all(phys(il + S * m) == phys(il) + phys(S * m) and phys(iu - S * m) == phys(iu) - phys(S * m)
                           for m in range(4))


def _sort32_pair_model(keys, cap_words):
    n, nw = len(keys), (len(keys) + 1) // 2
    a = list(keys)

    def ld(w):
        if w >= nw:
            return [INF, INF]
        return [a[2 * w], a[2 * w + 1] if 2 * w + 1 < n else INF]

    def st(w, v):
        if w >= nw:
            assert v == [INF, INF]; return
        a[2 * w] = v[0]
        if 2 * w + 1 < n:
            a[2 * w + 1] = v[1]
        else:
            assert v[1] == INF

    def cew(W, i, j):
        for h in (0, 1):
            if W[i][h] > W[j][h]: W[i][h], W[j][h] = W[j][h], W[i][h]

    def inword(W, i):
        if W[i][0] > W[i][1]: W[i][0], W[i][1] = W[i][1], W[i][0]

    def flip(W, lo, hi):
        if W[lo][0] > W[hi][1]: W[lo][0], W[hi][1] = W[hi][1], W[lo][0]
        if W[lo][1] > W[hi][0]: W[lo][1], W[hi][0] = W[hi][0], W[lo][1]

    g = 0
    while g * 8 < nw:                   # element levels 2, 4, 8, 16 on 8 consecutive pair words
        W = [ld(8 * g + m) for m in range(8)]
        for i in range(8): inword(W, i)
        for p in [(0, 1), (2, 3), (4, 5), (6, 7)]: flip(W, *p)
        for i in range(8): inword(W, i)
        for p in [(0, 3), (1, 2), (4, 7), (5, 6)]: flip(W, *p)
        for p in [(0, 1), (2, 3), (4, 5), (6, 7)]: cew(W, *p)
        for i in range(8): inword(W, i)
        for p in [(0, 7), (1, 6), (2, 5), (3, 4)]: flip(W, *p)
        for p in [(0, 2), (1, 3), (4, 6), (5, 7)]: cew(W, *p)
        for p in [(0, 1), (2, 3), (4, 5), (6, 7)]: cew(W, *p)
        for i in range(8): inword(W, i)
        for m in range(8): st(8 * g + m, W[m])
        g += 1
    Kw = 16
    while Kw <= cap_words and Kw // 2 < nw:
        S = Kw // 8
        for g in range(((nw + Kw - 1) // Kw) * S):
            r, base = g & (S - 1), (g // S) * Kw
            il, iu = base + r, base + Kw - 1 - r
            if il >= nw:
                continue
            W = [ld(il + S * m) for m in range(4)] + [ld(iu - S * m) for m in range(4)]
            for m in range(4): flip(W, m, 4 + m)
            for p in [(0, 2), (1, 3), (6, 4), (7, 5), (0, 1), (2, 3), (5, 4), (7, 6)]: cew(W, *p)
            for m in range(4): st(il + S * m, W[m]); st(iu - S * m, W[4 + m])
        J = Kw // 16
        while J >= 1:
            cnt = min(3, _ilog2(J) + 1)
            S = J >> (cnt - 1)
            for g in range(((nw + 8 * S - 1) // (8 * S)) * S):
                i0 = (g // S) * 8 * S + (g & (S - 1))
                if i0 >= nw:
                    continue
                W = [ld(i0 + S * m) for m in range(8)]
                if cnt >= 3:
                    for p in [(0, 4), (1, 5), (2, 6), (3, 7)]: cew(W, *p)
                if cnt >= 2:
                    for p in [(0, 2), (1, 3), (4, 6), (5, 7)]: cew(W, *p)
                for p in [(0, 1), (2, 3), (4, 5), (6, 7)]: cew(W, *p)
                if S == 1:
                    for i in range(8): inword(W, i)
                for m in range(8): st(i0 + S * m, W[m])
            J = S // 2
        Kw *= 2
    return a


def test_pair_word_network_sorts_every_length():
    rng = random.Random(3)
    for n in list(range(1, 80)) + [127, 128, 129, 255, 256, 257, 511, 513, 1000, 1023, 1024, 1025, 2000, 2047]:
        a = [rng.randrange(0, 1 << 20) for _ in range(n)]
        assert _sort32_pair_model(a, 1024) == sorted(a), n


def test_transposing_reduction_of_ten_values():
    rng = random.Random(0)
    V = [[rng.random() for _ in range(10)] for _ in range(32)]
    shfl = lambda vals, d: [vals[l ^ d] for l in range(32)]
    w = [[0.0] * 5 for _ in range(32)]
    for i in range(5):
        r = shfl([V[l][i] if l & 16 else V[l][i + 5] for l in range(32)], 16)
        for l in range(32): w[l][i] = (V[l][i + 5] if l & 16 else V[l][i]) + r[l]
    x = [[0.0] * 3 for _ in range(32)]
    for k in range(3):
        r = shfl([w[l][k] if l & 8 else (w[l][3 + k] if k < 2 else 0.0) for l in range(32)], 8)
        for l in range(32): x[l][k] = ((w[l][3 + k] if k < 2 else 0.0) if l & 8 else w[l][k]) + r[l]
    r0 = shfl([x[l][0] if l & 4 else x[l][2] for l in range(32)], 4)
    r1 = shfl([x[l][1] if l & 4 else 0.0 for l in range(32)], 4)
    y = [[(x[l][2] if l & 4 else x[l][0]) + r0[l], (0.0 if l & 4 else x[l][1]) + r1[l]] for l in range(32)]
    r = shfl([y[l][0] if l & 2 else y[l][1] for l in range(32)], 2)
    z = [(y[l][1] if l & 2 else y[l][0]) + r[l] for l in range(32)]
    r = shfl(z, 1)
    z = [z[l] + r[l] for l in range(32)]

    def red10_index(lane):
        if lane & 1: return -1
        b3, b2, b1 = (lane >> 3) & 1, (lane >> 2) & 1, (lane >> 1) & 1
        k = ((-1 if b1 else 2) if b2 else b1) if not b3 else (-1 if b2 else 3 + b1)
        return -1 if k < 0 else 5 * ((lane >> 4) & 1) + k

    total = [sum(V[l][i] for l in range(32)) for i in range(10)]
    seen = set()
    for l in range(32):
        i = red10_index(l)
        if i >= 0:
            assert abs(z[l] - total[i]) < 1e-9
            seen.add(i)
    assert seen == set(range(10))
