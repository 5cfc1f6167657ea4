#!/usr/bin/env python3
"""Derive the tables of the latency-optimised lane-cooperative Poseidon-Goldilocks permutation
(stark-verifier_b200/csrc/poseidon_g_coop2.cuh) from the committed parameter set
(poseidon_g_constants.inc, itself extracted from the reference's chip/plonk/gates/poseidon.rs:26-322)
and check the re-formulation against the plain fast form of poseidon.rs:634-686 in big-integer arithmetic.

The 22 partial rounds are linear except for ONE x^7 per round, so every quantity of the partial section is a
linear form over  (1, z_0..z_11, q_0..q_21)  where z = the S-box outputs of full round 3 (before its MDS
layer) and q_r = 25 * x_r^7 (x_r = lane 0 entering partial round r; u_r = x_r^7 + c_r = q_r / 25 + c_r is
what the fast form multiplies by its v / w_hat vectors).  In particular
    x_{r+1} = q_r + B_{r+1}(1, z, q_0..q_{r-1}).
The factor 25 (MDS_MATRIX_CIRC[0] + MDS_MATRIX_DIAG[0]) is scaled away: with a_0 = 1, a_{r+1} = 25 a_r^7 and
x_r = a_r t_r the recurrence is
    t_{r+1} = t_r^7 + B'_{r+1}(1, z, p_0..p_{r-1}),     p_r = t_r^7 = q_r / a_{r+1},   B'_r = B_r / a_r,
so the only work between two S-boxes is the addend of the last multiplication: t_{r+1} = t_r^3 * t_r^4 + B'_{r+1}.
The B' rows (and the 12 rows of the state that leaves the partial section, with the constants of round 26
folded in) are spread over the 16 lanes of a group as three accumulator slots per lane:
    slot 0, lane l      : B'_l           (l = 0..15)
    slot 1, lane l      : B'_{16+l}      (l = 0..5)
    slot 2, lane l      : F_l            (l = 0..11; F_0 = x_22 + rc_26[0], F_i = s_i(22) + rc_26[i])
Tables (u64, lane-minor so that a group reads 16 consecutive words):
    COOP2_ZC[k][slot][lane]  coefficient of z_k          (12 x 3 x 16)
    COOP2_QC[r][slot][lane]  coefficient of p_{r-1}      (23 x 3 x 16; row 0 is zero)
    COOP2_C0[slot][lane]     constant term               (3 x 16)
"""
import pathlib, random, re

ROOT = pathlib.Path(__file__).resolve().parent.parent
INC = ROOT / "stark-verifier_b200" / "csrc" / "poseidon_g_constants.inc"
OUT = ROOT / "stark-verifier_b200" / "csrc" / "poseidon_g_coop2_constants.inc"
P = 2**64 - 2**32 + 1


def tables():
    t = INC.read_text()
    out = {}
    for m in re.finditer(r"SVB_TABLE\((\w+),\s*(\d+)\)\s*=\s*\{(.*?)\};", t, re.S):
        vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(3))]
        assert len(vals) == int(m.group(2)), m.group(1)
        out[m.group(1)] = vals
    return out


T = tables()
RC, FC, PRC = T["ALL_ROUND_CONSTANTS"], T["FAST_PARTIAL_FIRST_ROUND_CONSTANT"], T["FAST_PARTIAL_ROUND_CONSTANTS"]
VS, WH, INIT = T["FAST_PARTIAL_ROUND_VS"], T["FAST_PARTIAL_ROUND_W_HATS"], T["FAST_PARTIAL_ROUND_INITIAL_MATRIX"]
CIRC, DIAG = T["MDS_MATRIX_CIRC"], T["MDS_MATRIX_DIAG"]
M = [[(CIRC[(j - i) % 12] + (DIAG[i] if i == j else 0)) % P for j in range(12)] for i in range(12)]
NB = 1 + 12 + 22          # basis: 1, z_0..z_11, q_0..q_21
INV25 = pow(25, P - 2, P)


def mds(v):
    return [sum(M[i][j] * v[j] for j in range(12)) % P for i in range(12)]


def sbox(x):
    return pow(x, 7, P)


def permute_fast(s):
    """poseidon.rs:634-686 (fast form), the function every implementation in this repository must equal."""
    s = list(s)
    for r in range(4):
        s = mds([sbox((s[i] + RC[12 * r + i]) % P) for i in range(12)])
    s = [(s[i] + FC[i]) % P for i in range(12)]
    s = [s[0]] + [sum(INIT[(r - 1) * 11 + (c - 1)] * s[r] for r in range(1, 12)) % P for c in range(1, 12)]
    for r in range(22):
        u = (sbox(s[0]) + PRC[r]) % P
        d = (25 * u + sum(WH[r * 11 + i - 1] * s[i] for i in range(1, 12))) % P
        s = [d] + [(s[i] + VS[r * 11 + i - 1] * u) % P for i in range(1, 12)]
    for r in range(26, 30):
        s = mds([sbox((s[i] + RC[12 * r + i]) % P) for i in range(12)])
    return s


# ---- linear forms ------------------------------------------------------------------------------------
def vec(const=0):
    v = [0] * NB
    v[0] = const % P
    return v


def axpy(a, x, y):      # a*x + y
    return [(a * xi + yi) % P for xi, yi in zip(x, y)]


def derive():
    z = []
    for k in range(12):
        e = vec()
        e[1 + k] = 1
        z.append(e)
    w = [vec() for _ in range(12)]
    for i in range(12):
        for j in range(12):
            w[i] = axpy(M[i][j], z[j], w[i])
    y = [axpy(1, w[i], vec(FC[i])) for i in range(12)]
    s = [y[0]]
    for c in range(1, 12):
        acc = vec()
        for r in range(1, 12):
            acc = axpy(INIT[(r - 1) * 11 + (c - 1)], y[r], acc)
        s.append(acc)
    X = []
    for r in range(22):
        X.append(s[0])
        u = vec(PRC[r])
        u[13 + r] = INV25
        d = axpy(25, u, vec())
        for i in range(1, 12):
            d = axpy(WH[r * 11 + i - 1], s[i], d)
        s = [d] + [axpy(VS[r * 11 + i - 1], u, s[i]) for i in range(1, 12)]
    F = [axpy(1, s[i], vec(RC[12 * 26 + i])) for i in range(12)]
    # B rows: x_r without its q_{r-1} term (whose coefficient is exactly 1)
    B = []
    for r in range(22):
        b = list(X[r])
        if r:
            assert b[13 + r - 1] == 1, r
            b[13 + r - 1] = 0
        assert all(c == 0 for c in b[13 + max(r - 1, 0):]), r
        B.append(b)
    # scale: x_r = a_r t_r, q_j = a_{j+1} p_j
    a = [1]
    for r in range(22):
        a.append(25 * pow(a[r], 7, P) % P)

    def cols(row):
        return row[:13] + [row[13 + j] * a[j + 1] % P for j in range(22)]
    B = [[c * pow(a[r], P - 2, P) % P for c in cols(B[r])] for r in range(22)]
    F = [cols(f) for f in F]
    return B, F


def lane_tables(B, F):
    rows = [[None] * 16 for _ in range(3)]
    for l in range(16):
        rows[0][l] = B[l]
        rows[1][l] = B[16 + l] if 16 + l < 22 else vec()
        rows[2][l] = F[l] if l < 12 else vec()
    ZC = [[[rows[s][l][1 + k] for l in range(16)] for s in range(3)] for k in range(12)]
    QC = [[[0 if r == 0 else rows[s][l][13 + r - 1] for l in range(16)] for s in range(3)] for r in range(23)]
    C0 = [[rows[s][l][0] for l in range(16)] for s in range(3)]
    return ZC, QC, C0


def permute_coop2(s, ZC, QC, C0):
    """The schedule of poseidon_g_coop2 in big integers: same tables, same order of operations."""
    s = [(s[i] + RC[i]) % P for i in range(12)]
    for f in range(3):
        zz = [sbox(x) for x in s]
        nxt = RC[12 * (f + 1):12 * (f + 2)]
        s = [(a + b) % P for a, b in zip(mds(zz), nxt)]
    z = [sbox(x) for x in s]
    acc = [[C0[sl][l] for l in range(16)] for sl in range(3)]
    for k in range(12):                           # slot 0: complete before the loop
        for l in range(16):
            acc[0][l] = (acc[0][l] + ZC[k][0][l] * z[k]) % P
    x = acc[0][0]                                  # x_0, broadcast from lane 0
    qprev = 0
    for r in range(22):
        for sl in range(3):                        # part A: q_{r-1} into every accumulator (row 0 of QC is zero)
            for l in range(16):
                acc[sl][l] = (acc[sl][l] + QC[r][sl][l] * qprev) % P
        if r < 12:                                 # lazy z-parts of slots 1 and 2
            for sl in (1, 2):
                for l in range(16):
                    acc[sl][l] = (acc[sl][l] + ZC[r][sl][l] * z[r]) % P
        nb = r + 1
        bn = acc[nb // 16][nb % 16] if nb < 22 else 0
        q = sbox(x)                             # p_r = t_r^7
        xn = (q + bn) % P
        qprev = (xn - bn) % P
        assert qprev == q
        x = xn
    for l in range(16):
        acc[2][l] = (acc[2][l] + QC[22][2][l] * qprev) % P
    s = acc[2][:12]                                # constants of round 26 already inside
    for f in range(4, 8):
        zz = [sbox(v) for v in s]
        nxt = RC[12 * (23 + f):12 * (24 + f)] if f < 7 else [0] * 12
        s = [(a + b) % P for a, b in zip(mds(zz), nxt)]
    return s


def main():
    assert permute_fast(list(range(12)))[0] == 0xd64e1e3efc5b8e9e       # SURVEY 8c known answer
    B, F = derive()
    ZC, QC, C0 = lane_tables(B, F)
    rnd = random.Random(5)
    cases = [list(range(12)), [0] * 12, [P - 1] * 12] + [[rnd.randrange(P) for _ in range(12)] for _ in range(40)]
    for s in cases:
        assert permute_coop2(s, ZC, QC, C0) == permute_fast(s)
    out = ["/* GENERATED by tools/gen_poseidon_coop2_constants.py from poseidon_g_constants.inc -- linear forms of the partial",
           " * section of Poseidon-Goldilocks over (1, z, q) for the lane-cooperative permutation; checked there against the fast",
           " * form in big-integer arithmetic.  Do not edit. */"]

    def emit(name, flat):
        out.append(f"SVB_TABLE({name}, {len(flat)}) = {{")
        for i in range(0, len(flat), 4):
            out.append("    " + ", ".join(f"0x{x:016x}ULL" for x in flat[i:i + 4]) + ",")
        out.append("};")
        out.append("")

    emit("COOP2_ZC", [ZC[k][s][l] for k in range(12) for s in range(3) for l in range(16)])
    emit("COOP2_QC", [QC[r][s][l] for r in range(23) for s in range(3) for l in range(16)])
    emit("COOP2_C0", [C0[s][l] for s in range(3) for l in range(16)])
    OUT.write_text("\n".join(out) + "\n")
    print("ok", OUT)


if __name__ == "__main__":
    main()
