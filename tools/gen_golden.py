#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

The reference is Rust and ships no golden vectors, so the fixtures are anchored as
follows:
  * poseidon_g.json -- the three SURVEY 8c known-answer permutations (zero / iota / p-1 states; the
    first two equal upstream plonky2's published test_vectors12) plus permutation, sponge and
    compression outputs on seeded inputs, produced by a PURE-PYTHON big-integer evaluation of the
    NAIVE 30-round Poseidon definition written in this script, reading the constants straight from
    the reference source (chip/plonk/gates/poseidon.rs:26-124, :321-322) when /root/reference exists
    (else from the generated oracle table).  It shares no code with oracle/oracle.c or the CUDA path.
  * fri_small.npz -- a handful of tiny-shape proof records (valid and corrupted) with the accept bit
    and first-failure code the oracle gives; regenerating needs the product's synthetic prover.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
P = 0xFFFFFFFF00000001
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
DIAG = [8] + [0] * 11


def round_constants():
    ref = "/root/reference/src/plonky2_verifier/chip/plonk/gates/poseidon.rs"
    if os.path.exists(ref):
        t = open(ref).read()
        body = re.search(r"const\s+ALL_ROUND_CONSTANTS\s*:[^=]*=\s*\[(.*?)\];", t, re.S).group(1)
    else:
        t = open(os.path.join(ROOT, "oracle", "poseidon_g_constants.h")).read()
        body = re.search(r"ORC_ALL_ROUND_CONSTANTS\[360\] = \{(.*?)\};", t, re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    v = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", body)]
    assert len(v) == 360
    return v


RC = round_constants()


def mds(s):
    return [(sum(CIRC[i] * s[(i + r) % 12] for i in range(12)) + DIAG[r] * s[r]) % P for r in range(12)]


def poseidon_py(s):
    s = [x % P for x in s]
    for rnd in range(30):
        s = [(x + RC[12 * rnd + i]) % P for i, x in enumerate(s)]
        if rnd < 4 or rnd >= 26:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = mds(s)
    return s


def hash_no_pad_py(inp):
    st = [0] * 12
    for off in range(0, len(inp), 8):
        chunk = inp[off:off + 8]
        st[:len(chunk)] = chunk
        st = poseidon_py(st)
    return st[:4]


# ---- hash family B: Poseidon over BN254 Fr wrapped around 12 Goldilocks limbs (pure big integers) ----
R_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def b_constants():
    ref = "/root/reference/src/plonky2_verifier/bn245_poseidon/constants.rs"
    if os.path.exists(ref):
        t = open(ref).read()
        rc = [int(x, 16) for x in re.findall(r'"0x([0-9a-fA-F]+)"', re.search(r"ROUND_CONSTANTS_STR[^=]*=\s*\[(.*?)\];", t, re.S).group(1))]
        mds = [int(x, 16) for x in re.findall(r'"0x([0-9a-fA-F]+)"', re.search(r"MDS_MATRIX_STR[^=]*=\s*\[(.*?)\];\s*\n\s*fn ", t, re.S).group(1))]
    else:
        t = open(os.path.join(ROOT, "oracle", "poseidon_b_constants.h")).read()
        def tab(name):
            w = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ULL", re.search(name + r"\[[^\]]*\] = \{(.*?)\};", t, re.S).group(1))]
            return [sum(w[4 * i + k] << (64 * k) for k in range(4)) for i in range(len(w) // 4)]
        rc, mds = tab("ORC_B_ROUND_CONSTANTS"), tab("ORC_B_MDS")
    assert len(rc) == 340 and len(mds) == 25
    return rc, [mds[5 * i:5 * i + 5] for i in range(5)]


def poseidon_b_fr_py(st, consts):
    """bn245_poseidon/native.rs:43-60 on Python integers."""
    rc, mds = consts
    st = list(st)
    c = 0
    for rnd in range(68):
        st = [(x + rc[c + i]) % R_BN for i, x in enumerate(st)]
        c += 5
        if rnd < 4 or rnd >= 64:
            st = [pow(x, 5, R_BN) for x in st]
        else:
            st[0] = pow(st[0], 5, R_BN)
        st = [sum(st[j] * mds[i][j] for j in range(5)) % R_BN for i in range(5)]
    return st


def poseidon_b_py(s12, consts):
    """Bn254PoseidonPermutation::permute (plonky2_config.rs:38-51) with encode_fe / decode_fe (native.rs:62-77)."""
    enc = [s12[3 * k] + s12[3 * k + 1] * P + s12[3 * k + 2] * P * P for k in range(4)] + [0]
    out = poseidon_b_fr_py(enc, consts)
    res = []
    for k in range(4):
        v = out[k]
        for _ in range(3):
            res.append(v % P)
            v //= P
    return res


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(0x601D)
    perm = []
    for s in ([0] * 12, list(range(12)), [P - 1] * 12):
        perm.append({"in": [f"{x:016x}" for x in s], "out": [f"{x:016x}" for x in poseidon_py(s)]})
    assert perm[0]["out"][0] == "3c18a9786cb0b359" and perm[1]["out"][0] == "d64e1e3efc5b8e9e" and perm[2]["out"][0] == "be0085cfc57a8357"
    for _ in range(13):
        s = [int(x) for x in rng.integers(0, P, size=12, dtype=np.uint64)]
        perm.append({"in": [f"{x:016x}" for x in s], "out": [f"{x:016x}" for x in poseidon_py(s)]})
    hashes = []
    for n in (1, 4, 5, 8, 9, 16, 20, 84, 135, 139):
        s = [int(x) for x in rng.integers(0, P, size=n, dtype=np.uint64)]
        hashes.append({"in": [f"{x:016x}" for x in s], "out": [f"{x:016x}" for x in hash_no_pad_py(s)]})
    json.dump({"source": "tools/gen_golden.py: pure-Python naive Poseidon over the reference's ALL_ROUND_CONSTANTS",
               "permutation": perm, "hash_no_pad": hashes}, open(os.path.join(out_dir, "poseidon_g.json"), "w"), indent=0)

    consts = b_constants()
    kat = poseidon_b_fr_py([0, 1, 2, 3, 4], consts)
    assert kat[0] == 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465   # circomlib poseidon([1,2,3,4])
    fr_cases = [{"in": [f"{x:064x}" for x in st], "out": [f"{x:064x}" for x in poseidon_b_fr_py(st, consts)]}
                for st in ([0, 1, 2, 3, 4], [0] * 5, [R_BN - 1] * 5,
                           [int.from_bytes(rng.bytes(32), "little") % R_BN for _ in range(5)])]
    wrapped = []
    for s12 in ([0] * 12, list(range(12)), [P - 1] * 12):
        wrapped.append({"in": [f"{x:016x}" for x in s12], "out": [f"{x:016x}" for x in poseidon_b_py(s12, consts)]})
    assert wrapped[1]["out"][0] == "d983775ce161c4e4" and wrapped[1]["out"][11] == "6bf843b27c9d3fbb"
    for _ in range(9):
        s12 = [int(x) for x in rng.integers(0, P, size=12, dtype=np.uint64)]
        wrapped.append({"in": [f"{x:016x}" for x in s12], "out": [f"{x:016x}" for x in poseidon_b_py(s12, consts)]})
    json.dump({"source": "tools/gen_golden.py: pure-Python Poseidon-BN254 (T=5, x^5, 8+60 rounds) over the reference's "
                         "bn245_poseidon/constants.rs, with the 3-limb base-p packing of native.rs:62-77",
               "fr_permutation": fr_cases, "wrapped_permutation": wrapped},
              open(os.path.join(out_dir, "poseidon_b.json"), "w"), indent=0)

    # FRI fixtures: tiny shapes, valid + corrupted, labelled by the oracle
    import stark_verifier_b200 as svb
    from oracle import binding as orc
    from common import corrupt, tiny_params
    fx = {}
    for tag, kw in (("plain", dict()), ("salted", dict(hiding=True, cap=1, degree_bits=6, rate_bits=2))):
        params = tiny_params(svb, **kw)
        L = svb.api.make_layout(params)
        recs = svb.synth_proofs(params, 18, seed=0x601D, n_circuits=2)
        corrupt(recs, L, np.random.default_rng(9), every=2)
        oshape = orc.shape_from(params.to_shape())
        res = [orc.fri_verify(oshape, recs[i]) for i in range(recs.shape[0])]
        fx[tag + "_records"] = recs
        fx[tag + "_accept"] = np.array([int(r[0]) for r in res], dtype=np.uint8)
        fx[tag + "_fail"] = np.array([0 if r[0] else ((max(r[2], 0) << 8) | r[1]) for r in res], dtype=np.uint32)
        s = params.to_shape()
        fx[tag + "_shape"] = np.array([s.degree_bits, s.rate_bits, s.cap_height, s.num_query_rounds, s.proof_of_work_bits,
                                       s.num_steps, s.final_poly_len, s.hiding], dtype=np.uint32)
    np.savez_compressed(os.path.join(out_dir, "fri_small.npz"), **fx)
    # one full-shape proof per BASELINE shape (A: configs[1], B: configs[2]); the shape-B prover run takes
    # about a minute on the host, which is why the proof is a committed fixture
    big = {}
    if os.path.exists(os.path.join(out_dir, "fri_full_shapes.npz")) and "--full" not in sys.argv:
        print("fri_full_shapes.npz kept (pass --full to regenerate: ~3 minutes)")
        return
    for tag, params in (("shape_a", svb.SHAPE_A), ("shape_b", svb.SHAPE_B)):
        rec = svb.synth_proofs(params, 1, seed=0xB2000003, n_circuits=1)
        ok, code, q = orc.fri_verify(orc.shape_from(params.to_shape()), rec[0])
        assert ok, (tag, code, q)
        cd, ph = svb.synth_public_inputs(params, 1, seed=0xB2000003, n_circuits=1)
        big[tag + "_record"] = rec[0]
        big[tag + "_circuit_digest"] = cd[0]
        big[tag + "_pi_hash"] = ph[0]
    np.savez(os.path.join(out_dir, "fri_full_shapes.npz"), **big)
    print("wrote", os.listdir(out_dir))


if __name__ == "__main__":
    main()
