#!/usr/bin/env python3
"""Index / twiddle model of the radix-8 register-blocked NTT passes (csrc/ntt_kernels.cuh), in plain Python integers:
the tile <-> global index map, the round structure (r <= 3 stages per round, top bits first), the twiddle factorisation
(one table entry w1 = omega_n^(u << s) per group, w1^bitrev(p) on output p) -- checked against the definition
(value at omega^bitrev(i) in position i).  Run: python tools/lab/ntt_model.py"""
import random

P = 0xFFFFFFFF00000001


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan(k, TL):
    """list of (s0, R): passes of a size-2^k DIF; non-final passes keep C = 2^(TL-R) >= 8 columns"""
    if k <= TL:
        return [(0, k)]
    last = TL - 1
    rest = k - last
    n_first = (rest + (TL - 3) - 1) // (TL - 3)
    out, s0 = [], 0
    for p in range(n_first):
        R = (rest - s0 + (n_first - p) - 1) // (n_first - p)
        out.append((s0, R))
        s0 += R
    out.append((s0, last))
    return out


def dif_pass(data, k, s0, R, TL, inverse=False):
    n = 1 << k
    w = pow(7, (P - 1) >> k, P)
    if inverse:
        w = pow(w, P - 2, P)
    lo_bits = k - s0 - R
    T = 1 << min(TL, k) if lo_bits else 1 << min(TL, k)
    T = min(T, n)
    logT = T.bit_length() - 1
    if lo_bits:      # m above the columns
        logC = logT - R
        mb = logC
    else:            # m lowest, consecutive hi blocks above it
        logC = logT - R
        mb = 0
    n_tiles = n // T
    for tile in range(n_tiles):
        # global index of tile element e
        if lo_bits:
            lo_blocks = 1 << (lo_bits - logC)
            hi, lob = divmod(tile, lo_blocks)
            base = (hi << (k - s0)) + (lob << logC)
        else:
            base = tile * T

        def gidx(e):
            return base + (e & ((1 << mb) - 1)) + ((e >> mb) << lo_bits) if lo_bits else base + e
        sm = [data[gidx(e)] for e in range(T)]
        # rounds: top m bits first (DIF); inverse: the same rounds backwards
        rounds = []
        top = R
        while top > 0:
            r = min(3, top) if top % 3 == 0 or top > 3 else top
            r = 3 if top >= 3 else top
            rounds.append((top - r, r))
            top -= r
        if inverse:
            rounds.reverse()
        for (b, r) in rounds:
            s_a = s0 + (R - b - r)            # first global stage of this round
            NA = n >> s_a                     # sub-transform size at that stage
            pos = mb + b                      # bit position of the round's bits inside e
            for g in range(T >> r):
                e0 = ((g >> pos) << (pos + r)) | (g & ((1 << pos) - 1))
                gi = gidx(e0)
                u = gi & ((NA >> r) - 1)      # index below the round's bits
                w1 = pow(w, u << s_a, P)      # omega_NA^u
                x = [sm[e0 | (j << pos)] for j in range(1 << r)]
                wr = pow(w, n >> r, P)        # primitive 2^r-th root (omega_8 for r = 3)
                if not inverse:
                    # r DIF stages inside the group, internal twiddles = powers of wr
                    size = 1 << r
                    st = 0
                    while size > 1:
                        half = size // 2
                        for blk in range(0, 1 << r, size):
                            for j in range(half):
                                a, c = x[blk + j], x[blk + j + half]
                                x[blk + j] = (a + c) % P
                                x[blk + j + half] = (a - c) * pow(wr, j << st, P) % P
                        size = half
                        st += 1
                    for p in range(1 << r):
                        x[p] = x[p] * pow(w1, bitrev(p, r), P) % P
                else:
                    for p in range(1 << r):
                        x[p] = x[p] * pow(w1, bitrev(p, r), P) % P
                    size = 2
                    st = r - 1
                    while size <= 1 << r:
                        half = size // 2
                        for blk in range(0, 1 << r, size):
                            for j in range(half):
                                a, c = x[blk + j], x[blk + j + half] * pow(wr, j << st, P) % P
                                x[blk + j] = (a + c) % P
                                x[blk + j + half] = (a - c) % P
                        size *= 2
                        st -= 1
                for j in range(1 << r):
                    sm[e0 | (j << pos)] = x[j]
        for e in range(T):
            data[gidx(e)] = sm[e]


def ntt_model(coeffs, k, TL, inverse=False):
    data = list(coeffs)
    passes = plan(k, TL)
    if inverse:
        passes = passes[::-1]
    for (s0, R) in passes:
        dif_pass(data, k, s0, R, TL, inverse)
    if inverse:
        ninv = pow(1 << k, P - 2, P)
        data = [v * ninv % P for v in data]
    return data


def main():
    rnd = random.Random(1)
    for k, TL in ((3, 13), (5, 13), (6, 5), (7, 5), (9, 6), (10, 7), (11, 6), (12, 7), (8, 13), (13, 7)):
        n = 1 << k
        c = [rnd.randrange(P) for _ in range(n)]
        w = pow(7, (P - 1) >> k, P)
        want = [sum(c[j] * pow(w, bitrev(i, k) * j, P) for j in range(n)) % P for i in range(n)] if k <= 10 else None
        got = ntt_model(c, k, TL)
        if want is not None:
            assert got == want, (k, TL)
        back = ntt_model(got, k, TL, inverse=True)
        assert back == c, (k, TL, "inverse")
        print("ok", k, TL, plan(k, TL))


if __name__ == "__main__":
    main()
