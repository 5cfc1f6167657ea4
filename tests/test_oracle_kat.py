"""Pin the oracle: known-answer vectors of SURVEY.md section 8c (computed from the reference's own
constants; the first two Poseidon vectors equal upstream plonky2's published test_vectors12), field
facts, and self-consistency of the two Poseidon forms."""
import json
import os

import numpy as np

from common import P

KAT = {
    "zero": ([0] * 12, "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc 46a84e87642f44ed "
                       "d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae d070f637b431067c 1792b1c4342109d7"),
    "iota": (list(range(12)), "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 cb1f8980294a023f "
                              "6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 c9bfd01d39b50cce 5c0a27fcb0e1459b"),
    "neg1": ([P - 1] * 12, "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 e095905e913a3029 "
                           "de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 e18263681e9958f8 e28e96f1ae5e60d3"),
}


def test_poseidon_kats(orc):
    for name, (inp, out) in KAT.items():
        want = [int(x, 16) for x in out.split()]
        assert [int(x) for x in orc.poseidon(inp)] == want, name
        assert [int(x) for x in orc.poseidon(inp, naive=True)] == want, name


def test_poseidon_fast_equals_naive_random(orc):
    rng = np.random.default_rng(1)
    for _ in range(200):
        s = rng.integers(0, P, size=12, dtype=np.uint64)
        assert (orc.poseidon(s) == orc.poseidon(s, naive=True)).all()


def test_tables_canonical(orc):
    assert orc.lib().orc_tables_canonical() == 1


def test_field_kats(orc):
    L = orc.lib()
    assert L.orc_f_pow(7, (P - 1) >> 32) == 1753635133440165772
    assert L.orc_f_pow(7, (P - 1) // 2) == P - 1
    # 7 generates F_p^*: 7^((p-1)/q) != 1 for every prime q | p-1 = 2^32 * 3 * 5 * 17 * 257 * 65537
    for q in (2, 3, 5, 17, 257, 65537):
        assert L.orc_f_pow(7, (P - 1) // q) != 1
    rng = np.random.default_rng(2)
    for _ in range(500):
        a, b = (int(x) for x in rng.integers(1, P, size=2, dtype=np.uint64))
        assert L.orc_f_mul(a, b) == (a * b) % P
        assert L.orc_f_mul(a, L.orc_f_inv(a)) == 1
        lo, hi = (int(x) for x in rng.integers(0, 1 << 63, size=2, dtype=np.uint64))
        lo, hi = lo * 2 + 1, hi * 2 + 1
        assert L.orc_f_red128(lo, hi, 0) == L.orc_f_red128(lo, hi, 1) == ((hi << 64) | lo) % P
    for lo, hi in ((0, 0), (2**64 - 1, 2**64 - 1), (0, 2**64 - 1), (2**64 - 1, 0), (P, P), (0, 2**32), (1, 0xFFFFFFFF)):
        assert L.orc_f_red128(lo, hi, 0) == ((hi << 64) | lo) % P


def test_ext_field(orc):
    L = orc.lib()
    rng = np.random.default_rng(3)
    for _ in range(100):
        a = rng.integers(0, P, size=2, dtype=np.uint64)
        b = rng.integers(0, P, size=2, dtype=np.uint64)
        out = np.zeros(2, dtype=np.uint64)
        L.orc_f2_mul(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        a0, a1, b0, b1 = int(a[0]), int(a[1]), int(b[0]), int(b[1])
        assert int(out[0]) == (a0 * b0 + 7 * a1 * b1) % P and int(out[1]) == (a0 * b1 + a1 * b0) % P
        inv = np.zeros(2, dtype=np.uint64)
        L.orc_f2_inv(a.ctypes.data, inv.ctypes.data)
        L.orc_f2_mul(a.ctypes.data, inv.ctypes.data, out.ctypes.data)
        assert int(out[0]) == 1 and int(out[1]) == 0


def test_sponge_and_compress(orc):
    rng = np.random.default_rng(4)
    x = rng.integers(0, P, size=135, dtype=np.uint64)
    # overwrite-mode sponge: <= 8 inputs is one permutation of [inputs, 0...]
    for n in (1, 5, 8):
        st = np.zeros(12, dtype=np.uint64)
        st[:n] = x[:n]
        assert (orc.hash_no_pad(x[:n]) == orc.poseidon(st)[:4]).all()
    # 9 inputs: second block overwrites lane 0 only
    st = np.zeros(12, dtype=np.uint64); st[:8] = x[:8]
    st = orc.poseidon(st); st[0] = x[8]
    assert (orc.hash_no_pad(x[:9]) == orc.poseidon(st)[:4]).all()
    l, r = x[:4], x[4:8]
    st = np.zeros(12, dtype=np.uint64); st[:4] = l; st[4:8] = r
    assert (orc.two_to_one(l, r) == orc.poseidon(st)[:4]).all()


def test_golden_fixtures_oracle(orc):
    """The committed fixtures (tests/golden/, made by tools/gen_golden.py) against the oracle."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "poseidon_g.json")))
    for v in g["permutation"]:
        assert [f"{int(x):016x}" for x in orc.poseidon([int(a, 16) for a in v["in"]])] == v["out"]
    for v in g["hash_no_pad"]:
        assert [f"{int(x):016x}" for x in orc.hash_no_pad(np.array([int(a, 16) for a in v["in"]], dtype=np.uint64))] == v["out"]


def test_oracle_accepts_full_shape_golden_proofs(orc):
    """tests/golden/fri_full_shapes.npz: one proof per BASELINE shape (A: 2^12 trace / 28 queries / blowup 8,
    B: 2^20 / 84 / 4).  The oracle accepts them, rejects a flipped bit, and its transcript reproduces the
    recorded challenges from (circuit digest, public-input hash)."""
    import ctypes
    import os
    import numpy as np
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "fri_full_shapes.npz"))
    for tag, (deg, rate, q) in (("shape_a", (12, 3, 28)), ("shape_b", (20, 2, 84))):
        s = orc.OrcShape()
        s.degree_bits, s.rate_bits, s.cap_height, s.num_query_rounds, s.proof_of_work_bits = deg, rate, 4, q, 16
        s.num_steps, s.final_poly_len, s.hiding = deg - 5, 32, 0
        for i in range(deg - 5):
            s.reduction_arity_bits[i] = 1            # ConstantArityBits(1, 5)
        s.oracle_num_polys = (ctypes.c_uint32 * 4)(84, 135, 20, 16)
        s.oracle_blinding = (ctypes.c_uint32 * 4)(0, 1, 1, 1)
        s.num_zs, s.hash_kind = 2, 0
        rec = np.ascontiguousarray(fx[tag + "_record"])
        ok, code, _ = orc.fri_verify(s, rec)
        assert ok and code == 0
        again = rec.copy()
        orc.fri_challenges(s, again, fx[tag + "_circuit_digest"], fx[tag + "_pi_hash"])
        assert (again == rec).all()
        bad = rec.copy()
        bad[len(bad) // 2] ^= np.uint64(1)
        assert not orc.fri_verify(s, bad)[0]


def test_oracle_poseidon_b_golden(orc):
    """Hash family B (Poseidon over BN254 Fr wrapped around 12 Goldilocks limbs): the oracle's Montgomery
    implementation against tests/golden/poseidon_b.json (pure-Python big integers over the reference's
    constants) and the SURVEY 8c KATs (state [0,1,2,3,4] -> circomlib poseidon([1,2,3,4]))."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "poseidon_b.json")))
    for c in d["fr_permutation"]:
        got = orc.poseidon_b_fr([int(x, 16) for x in c["in"]])
        assert got == [int(x, 16) for x in c["out"]]
    assert orc.poseidon_b_fr([0, 1, 2, 3, 4])[0] == 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465
    for c in d["wrapped_permutation"]:
        got = orc.poseidon_b([int(x, 16) for x in c["in"]])
        assert [int(x) for x in got] == [int(x, 16) for x in c["out"]]
    w = orc.poseidon_b(list(range(12)))
    assert int(w[0]) == 0xd983775ce161c4e4 and int(w[11]) == 0x6bf843b27c9d3fbb


def test_oracle_matches_upstream_plonky2_random_vector(orc):
    """The random-state vector of plonky2's own `test_vectors12` (fast and naive forms)."""
    from common import PLONKY2_TV12_IN, PLONKY2_TV12_OUT
    assert [int(x) for x in orc.poseidon(PLONKY2_TV12_IN)] == PLONKY2_TV12_OUT
    assert [int(x) for x in orc.poseidon(PLONKY2_TV12_IN, naive=True)] == PLONKY2_TV12_OUT
