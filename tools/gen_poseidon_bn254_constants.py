#!/usr/bin/env python3
"""Extract the Poseidon-BN254 parameter set (hash family "B": T = 5, x^5, R_F = 8, R_P = 60) from the
reference and emit generated tables for the CPU oracle (canonical 4 x u64 limbs) and for the product
(Montgomery form, R = 2^256, 8 x u32 limbs = 4 x u64 words).

Source of the numbers: /root/reference/src/plonky2_verifier/bn245_poseidon/constants.rs:5-384
(ROUND_CONSTANTS_STR: 340 hex strings, MDS_MATRIX_STR: 5 x 5), :402-404 (T, R_F, R_P).  Only numeric
literals are read; no code is copied.  The generated files are committed (the reference does not exist
on the GPU box).
"""
import pathlib
import re

REF = pathlib.Path("/root/reference/src/plonky2_verifier/bn245_poseidon/constants.rs")
ROOT = pathlib.Path(__file__).resolve().parent.parent
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617   # BN254 scalar field
GL_P = 0xFFFFFFFF00000001


def load():
    t = REF.read_text()
    rc_body = re.search(r"ROUND_CONSTANTS_STR[^=]*=\s*\[(.*?)\];", t, re.S).group(1)
    rc = [int(x, 16) for x in re.findall(r'"0x([0-9a-fA-F]+)"', rc_body)]
    mds_body = re.search(r"MDS_MATRIX_STR[^=]*=\s*\[(.*?)\];\s*\n\s*fn ", t, re.S).group(1)
    mds = [int(x, 16) for x in re.findall(r'"0x([0-9a-fA-F]+)"', mds_body)]
    assert len(rc) == 340 and len(mds) == 25, (len(rc), len(mds))
    assert all(v < R_MOD for v in rc + mds)
    return rc, mds


# ---- sparse form of the 60 partial rounds -------------------------------------------------------------
# Round r of the partial section is s <- M S(s + c_r), S = x^5 on lane 0 only.  A matrix D = [[1,0],[0,B]]
# commutes with S, and any M = [[m00, w^T],[v, Mh]] factors as Sp * D with Sp = [[m00, w^T Mh^-1],[v, I]],
# D = [[1,0],[0,Mh]].  Walking from the last partial round to the first and handing each D to the round
# before it (M_prev <- D M) leaves: one dense 4x4 block D_0 in front, then per round
#     t <- S(t);  t <- Sp_r t;  t <- t + D_{r+1} c_{r+1}   (the last addition only for r < 59)
# with t_0 = D_0 (s + c_0).  Sp_r t costs one 5-term dot product (lane 0) and four multiply-adds instead
# of five 5-term dot products.  The result is the same permutation (checked below against the naive form).
def mat_inv(a):
    n = len(a)
    m = [list(row) + [int(i == j) for j in range(n)] for i, row in enumerate(a)]
    for c in range(n):
        piv = next(r for r in range(c, n) if m[r][c] % R_MOD)
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], -1, R_MOD)
        m[c] = [x * inv % R_MOD for x in m[c]]
        for r in range(n):
            if r != c and m[r][c]:
                f = m[r][c]
                m[r] = [(x - f * y) % R_MOD for x, y in zip(m[r], m[c])]
    return [row[n:] for row in m]


def mat_mul(a, b):
    return [[sum(a[i][k] * b[k][j] for k in range(len(b))) % R_MOD for j in range(len(b[0]))] for i in range(len(a))]


def sparse_partial(rc, mds):
    M = [mds[5 * i:5 * i + 5] for i in range(5)]
    c = [rc[5 * (4 + r):5 * (4 + r) + 5] for r in range(60)]
    cur = M
    w_hat, v, cprime, D = [None] * 60, [None] * 60, [None] * 60, [None] * 60
    for r in range(59, -1, -1):
        Mh = [row[1:] for row in cur[1:]]
        Mh_inv = mat_inv(Mh)
        w = [cur[0][1:]]
        w_hat[r] = mat_mul(w, Mh_inv)[0]
        v[r] = [cur[i][0] for i in range(1, 5)]
        assert cur[0][0] == M[0][0]
        D[r] = [[1, 0, 0, 0, 0]] + [[0] + Mh[i] for i in range(4)]
        cprime[r] = [sum(D[r][i][k] * c[r][k] for k in range(5)) % R_MOD for i in range(5)]
        cur = mat_mul(D[r], M)
    return {"m00": M[0][0], "w_hat": w_hat, "v": v, "cprime": cprime, "D0": [row[1:] for row in D[0][1:]]}


def perm_naive(st, rc, mds):
    M = [mds[5 * i:5 * i + 5] for i in range(5)]
    k = 0
    for rnd in range(68):
        st = [(x + rc[k + i]) % R_MOD for i, x in enumerate(st)]
        k += 5
        st = [pow(x, 5, R_MOD) for x in st] if (rnd < 4 or rnd >= 64) else [pow(st[0], 5, R_MOD)] + st[1:]
        st = [sum(M[i][j] * st[j] for j in range(5)) % R_MOD for i in range(5)]
    return st


def perm_sparse(st, rc, mds, sp):
    M = [mds[5 * i:5 * i + 5] for i in range(5)]
    def full(st, rnd):
        st = [pow((x + rc[5 * rnd + i]) % R_MOD, 5, R_MOD) for i, x in enumerate(st)]
        return [sum(M[i][j] * st[j] for j in range(5)) % R_MOD for i in range(5)]
    for rnd in range(4):
        st = full(st, rnd)
    # t_0 = D_0 (s + c_0) = D_0 s + c'_0
    t = [st[0]] + [sum(sp["D0"][i][j] * st[1 + j] for j in range(4)) % R_MOD for i in range(4)]
    t = [(x + y) % R_MOD for x, y in zip(t, sp["cprime"][0])]
    for r in range(60):
        x = pow(t[0], 5, R_MOD)
        n0 = (sp["m00"] * x + sum(sp["w_hat"][r][j] * t[1 + j] for j in range(4))) % R_MOD
        t = [n0] + [(t[1 + j] + sp["v"][r][j] * x) % R_MOD for j in range(4)]
        if r < 59:
            t = [(a + b) % R_MOD for a, b in zip(t, sp["cprime"][r + 1])]
    st = t
    for rnd in range(64, 68):
        st = full(st, rnd)
    return st


def limbs64(v):
    return [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]


def emit_table(out, decl, vals):
    out.append(decl + " = {")
    for v in vals:
        out.append("    " + ", ".join(f"0x{x:016x}ULL" for x in limbs64(v)) + ",")
    out.append("};")
    out.append("")


def main():
    rc, mds = load()
    R = (1 << 256) % R_MOD
    hdr = ["/* GENERATED by tools/gen_poseidon_bn254_constants.py -- Poseidon over BN254 Fr, T = 5, x^5,",
           " * R_F = 8, R_P = 60.  Data source: reference bn245_poseidon/constants.rs:5-384,402-404.  Do not edit.",
           " * Every value is 4 little-endian u64 limbs. */"]
    # oracle: canonical values
    o = hdr + ["#ifndef ORC_POSEIDON_B_CONSTANTS_H", "#define ORC_POSEIDON_B_CONSTANTS_H", "#include <stdint.h>", ""]
    emit_table(o, "static const uint64_t ORC_B_ROUND_CONSTANTS[340 * 4]", rc)
    emit_table(o, "static const uint64_t ORC_B_MDS[25 * 4]", mds)
    o.append("#endif")
    (ROOT / "oracle" / "poseidon_b_constants.h").write_text("\n".join(o) + "\n")
    # product: Montgomery form, X-macro like poseidon_g_constants.inc
    x = hdr[:-1] + [" * Every value is 4 little-endian u64 limbs, in MONTGOMERY form (v * 2^256 mod r).",
                    " * Usage: #define SVB_TABLE(name, n) <qualifiers> uint64_t <prefix>##name[n]  then include. */"]
    emit_table(x, "SVB_TABLE(B_ROUND_CONSTANTS_MONT, 340 * 4)", [v * R % R_MOD for v in rc])
    emit_table(x, "SVB_TABLE(B_MDS_MONT, 25 * 4)", [v * R % R_MOD for v in mds])
    # sparse partial rounds (derived; verified against the naive permutation right here)
    import random
    sp = sparse_partial(rc, mds)
    rnd = random.Random(254)
    for _ in range(5):
        st = [rnd.randrange(R_MOD) for _ in range(5)]
        assert perm_naive(st, rc, mds) == perm_sparse(st, rc, mds, sp)
    mont = lambda v: v * R % R_MOD
    # per partial round r: [m00, w_hat_r[0..3]] (the lane-0 row), v_r[0..3], c'_r[0..4]  => 14 elements
    per_round = []
    for r in range(60):
        per_round += [mont(sp["m00"])] + [mont(x) for x in sp["w_hat"][r]] + [mont(x) for x in sp["v"][r]] + [mont(x) for x in sp["cprime"][r]]
    emit_table(x, "SVB_TABLE(B_SPARSE_ROUNDS_MONT, 60 * 14 * 4)", per_round)
    emit_table(x, "SVB_TABLE(B_SPARSE_D0_MONT, 16 * 4)", [mont(sp["D0"][i][j]) for i in range(4) for j in range(4)])
    (ROOT / "stark-verifier_b200" / "csrc" / "poseidon_b_constants.inc").write_text("\n".join(x) + "\n")
    print("ok: r =", hex(R_MOD), " -r^-1 mod 2^32 =", hex((-pow(R_MOD, -1, 1 << 32)) % (1 << 32)),
          " -r^-1 mod 2^64 =", hex((-pow(R_MOD, -1, 1 << 64)) % (1 << 64)))
    print("R mod r =", [hex(v) for v in limbs64(R)])
    print("R^2 mod r =", [hex(v) for v in limbs64(R * R % R_MOD)])
    print("p^-1 mod 2^256 =", [hex(v) for v in limbs64(pow(GL_P, -1, 1 << 256))])


if __name__ == "__main__":
    main()
