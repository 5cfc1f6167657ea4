#!/usr/bin/env python3
"""Extract the Poseidon-Goldilocks parameter set (pure data) from the reference and emit
two generated tables: one for the CPU oracle (C) and one for the product (C++/CUDA).

Source of the numbers: /root/reference/src/plonky2_verifier/chip/plonk/gates/poseidon.rs:26-322
(ALL_ROUND_CONSTANTS, FAST_PARTIAL_*, MDS_MATRIX_CIRC/DIAG).  Only numeric literals are read;
no code is copied.  The generated files are committed, so this script is only needed when
re-deriving them (it needs /root/reference, which does not exist on the GPU box).
"""
import re, sys, pathlib

REF = pathlib.Path("/root/reference/src/plonky2_verifier/chip/plonk/gates/poseidon.rs")
ROOT = pathlib.Path(__file__).resolve().parent.parent

def grab(text, name):
    m = re.search(r"const\s+" + name + r"\s*:[^=]*=\s*\[(.*?)\];", text, re.S)
    assert m, name
    body = re.sub(r"//[^\n]*", "", m.group(1))
    vals = [int(x, 16) if x.lower().startswith("0x") else int(x)
            for x in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", body)]
    return vals

def main():
    t = REF.read_text()
    tabs = {
        "ALL_ROUND_CONSTANTS": (grab(t, "ALL_ROUND_CONSTANTS"), 360),
        "FAST_PARTIAL_FIRST_ROUND_CONSTANT": (grab(t, "FAST_PARTIAL_FIRST_ROUND_CONSTANT"), 12),
        "FAST_PARTIAL_ROUND_CONSTANTS": (grab(t, "FAST_PARTIAL_ROUND_CONSTANTS"), 22),
        "FAST_PARTIAL_ROUND_VS": (grab(t, "FAST_PARTIAL_ROUND_VS"), 22 * 11),
        "FAST_PARTIAL_ROUND_W_HATS": (grab(t, "FAST_PARTIAL_ROUND_W_HATS"), 22 * 11),
        "FAST_PARTIAL_ROUND_INITIAL_MATRIX": (grab(t, "FAST_PARTIAL_ROUND_INITIAL_MATRIX"), 121),
        "MDS_MATRIX_CIRC": (grab(t, "MDS_MATRIX_CIRC"), 12),
        "MDS_MATRIX_DIAG": (grab(t, "MDS_MATRIX_DIAG"), 12),
    }
    for k, (v, n) in tabs.items():
        assert len(v) == n, (k, len(v), n)
    # derived (product only): constants folded into the MDS layer that follows full round f
    rc = tabs["ALL_ROUND_CONSTANTS"][0]
    nxt = rc[12:48] + tabs["FAST_PARTIAL_FIRST_ROUND_CONSTANT"][0] + rc[27 * 12:30 * 12] + [0] * 12
    # the same constants as binary64 bit patterns of 2^52 + low half, 2^52 + high half: the starting
    # values of the exact-integer FP64 accumulators of the MDS layer (poseidon_g.cuh)
    f64 = []
    for x in nxt:
        f64 += [0x4330000000000000 | (x & 0xFFFFFFFF), 0x4330000000000000 | (x >> 32)]
    # and as SUBNORMAL binary64 bit patterns (low half, high half as plain 64-bit integers): the starting
    # values of the accumulators when the MDS inputs are fed as subnormals (no int->double conversion)
    sub = []
    for x in nxt:
        sub += [x & 0xFFFFFFFF, x >> 32]
    # ---- naive-round schedule of the device path (poseidon_g.cuh, SVB_PARTIAL_NAIVE) ------------------------
    # Rounds 4..25 apply x^7 to lane 0 only, so the constants of lanes 1..11 commute with the S-box layer and
    # can be pushed through the (linear) MDS into the next round: with rest_r = d_r minus its lane 0,
    #   d_5 = c_5,  d_{r+1} = c_{r+1} + M * rest_r  (r = 5..25),
    # every partial round adds ONE scalar d_r[0] to lane 0 and the whole of d_26 is added before round 26.
    P = 2**64 - 2**32 + 1
    circ, diag = tabs["MDS_MATRIX_CIRC"][0], tabs["MDS_MATRIX_DIAG"][0]
    def mds(v):
        return [(sum(circ[k] * v[(i + k) % 12] for k in range(12)) + diag[i] * v[i]) % P for i in range(12)]
    c = [rc[12 * r:12 * r + 12] for r in range(30)]
    d = {5: c[5]}
    for r in range(5, 26):
        rest = [0] + d[r][1:]
        d[r + 1] = [(x + y) % P for x, y in zip(c[r + 1], mds(rest))]
    def halves(v):
        out = []
        for x in v:
            out += [x & 0xFFFFFFFF, x >> 32]
        return out
    # MDS layer after full round f = 0..7 (rounds 0..3, 26..29) adds the constants of the next round
    # (after round 3: c_4, complete; after round 29: nothing), as subnormal halves
    naive_full = halves(c[1] + c[2] + c[3] + c[4] + c[27] + c[28] + c[29] + [0] * 12)
    # MDS layer after partial round r = 4..25 adds d_{r+1}[0] to lane 0 (nothing after round 25)
    naive_lane0 = halves([d[r + 1][0] for r in range(4, 25)] + [0])
    # The same constants moved IN FRONT of that MDS layer: M^-1 * c_next, added to the S-box outputs of full round f, where
    # the addition rides on the last multiplication of x^7 (x3 * x4 + c) instead of costing 24 FP64 additions per layer.
    def mds_inverse_apply(v):
        n = 12
        A = [[(circ[(j - i) % 12] + (diag[i] if i == j else 0)) % P for j in range(n)] + [v[i] % P] for i in range(n)]
        for col in range(n):
            piv = next(r for r in range(col, n) if A[r][col])
            A[col], A[piv] = A[piv], A[col]
            inv = pow(A[col][col], P - 2, P)
            A[col] = [x * inv % P for x in A[col]]
            for r in range(n):
                if r != col and A[r][col]:
                    f = A[r][col]
                    A[r] = [(x - f * y) % P for x, y in zip(A[r], A[col])]
        return [A[i][n] for i in range(n)]
    pre_mds = []
    for v in (c[1], c[2], c[3], c[4], c[27], c[28], c[29], [0] * 12):
        w = mds_inverse_apply(v)
        assert mds(w) == [x % P for x in v]
        pre_mds += w
    derived = {"NAIVE_FULL_RC_PRE_MDS": (pre_mds, 96), "FULL_RC_NEXT": (nxt, 96), "FULL_RC_NEXT_F64": (f64, 192), "FULL_RC_NEXT_SUBNORMAL": (sub, 192),
               "NAIVE_FULL_RC_NEXT_SUBNORMAL": (naive_full, 192), "NAIVE_LANE0_RC_SUBNORMAL": (naive_lane0, 44),
               "NAIVE_PRE_ROUND26": (d[26], 12)}

    def emit_c(path, prefix, guard):
        out = [f"/* GENERATED by tools/gen_poseidon_constants.py -- Poseidon-Goldilocks parameters",
               f" * (width 12, x^7, 8 full + 22 partial rounds).  Data source:",
               f" * reference chip/plonk/gates/poseidon.rs:26-322.  Do not edit. */",
               f"#ifndef {guard}", f"#define {guard}", "#include <stdint.h>", ""]
        for k, (v, n) in tabs.items():
            out.append(f"static const uint64_t {prefix}{k}[{n}] = {{")
            for i in range(0, n, 4):
                out.append("    " + ", ".join(f"0x{x:016x}ULL" for x in v[i:i+4]) + ",")
            out.append("};")
            out.append("")
        out.append(f"#endif /* {guard} */")
        path.write_text("\n".join(out) + "\n")

    def emit_x(path):
        # X-macro form: the includer defines SVB_TABLE(name, n) (e.g. once as a host table and
        # once as a __constant__ table) and includes this file.
        out = ["/* GENERATED by tools/gen_poseidon_constants.py -- Poseidon-Goldilocks parameters",
               " * (width 12, x^7, 8 full + 22 partial rounds).  Data source:",
               " * reference chip/plonk/gates/poseidon.rs:26-322.  Do not edit.",
               " * Usage: #define SVB_TABLE(name, n) <qualifiers> uint64_t <prefix>##name[n]  then include. */"]
        for k, (v, n) in list(tabs.items()) + list(derived.items()):
            out.append(f"SVB_TABLE({k}, {n}) = {{")
            for i in range(0, n, 4):
                out.append("    " + ", ".join(f"0x{x:016x}ULL" for x in v[i:i+4]) + ",")
            out.append("};")
            out.append("")
        path.write_text("\n".join(out) + "\n")

    emit_c(ROOT / "oracle" / "poseidon_g_constants.h", "ORC_", "ORC_POSEIDON_G_CONSTANTS_H")
    emit_x(ROOT / "stark-verifier_b200" / "csrc" / "poseidon_g_constants.inc")
    print("ok")

if __name__ == "__main__":
    main()
