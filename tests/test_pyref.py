"""The independent pure-Python reference (tests/pyref/) against the C oracle and the product's host side, and the fixtures it
produced (tests/golden/pyref_fri.npz) against both.  What this pins: two restatements written separately from the reference's
Rust -- one in C (oracle/), one in Python (tests/pyref/, which also contains a PROVER) -- agree on every permutation, every
challenge, every wire byte and every verdict including the first-failure code; -m gpu repeats the verdicts on the CUDA path
(tests/test_gpu_pyref_golden.py)."""
import os
import re

import numpy as np
import pytest

import golden_pyref as gp
from common import P, PLONKY2_TV12_IN, PLONKY2_TV12_OUT, bit, corrupt
from pyref import challenger as pch
from pyref import fri as pfri
from pyref import gl
from pyref import poseidon as ps
from pyref import proof as ppf
from pyref import prover as ppr

HERE = os.path.dirname(os.path.abspath(__file__))


def rf(rng, lo=0):
    return int(rng.integers(lo, P, dtype=np.uint64))


def test_pyref_shares_no_code_with_oracle_or_product():
    for root in ("pyref",):
        for f in os.listdir(os.path.join(HERE, root)):
            if f.endswith(".py"):
                code = open(os.path.join(HERE, root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|stark_verifier_b200|ctypes|common|full_prover)", code, re.M), f
                assert "binding" not in code and "libsvb200" not in code and "liboracle" not in code, f
    fp = open(os.path.join(HERE, "full_prover.py")).read()
    assert not re.search(r"\borc\.", fp) and "oracle" not in fp.split('"""')[2]
    gen = open(os.path.join(os.path.dirname(HERE), "tools", "gen_golden_pyref.py")).read()
    assert not re.search(r"^\s*(from|import)\s+(oracle|stark_verifier_b200)", gen, re.M)


def test_field_vectors_match_the_scalar_definition():
    rng = np.random.default_rng(1)
    corner = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, 0xFFFFFFFE00000001, 1 << 63, (1 << 63) - 1,
              0xFFFFFFFEFFFFFFFF, 0x7FFFFFFF80000001]
    a = corner * len(corner) + [int(v) for v in rng.integers(0, P, size=5000, dtype=np.uint64)]
    b = [c for c in corner for _ in corner] + [int(v) for v in rng.integers(0, P, size=5000, dtype=np.uint64)]
    A, B = gl.varr(a), gl.varr(b)
    assert [int(x) for x in gl.vmul(A, B)] == [x * y % P for x, y in zip(a, b)]
    assert [int(x) for x in gl.vadd(A, B)] == [(x + y) % P for x, y in zip(a, b)]
    assert [int(x) for x in gl.vsub(A, B)] == [(x - y) % P for x, y in zip(a, b)]
    z = (rf(rng, 1), rf(rng, 1))
    assert gl.e_mul(z, gl.e_inv(z)) == (1, 0)
    assert gl.root_of_unity(32) == 1753635133440165772 and gl.root_of_unity(1) == P - 1     # SURVEY 8c field KATs


def test_poseidon_known_answers_and_oracle(orc):
    # SURVEY 8c KATs; the first two and the fourth are upstream plonky2's test_vectors12
    assert ps.permute_g([0] * 12)[:2] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3]
    assert ps.permute_g(list(range(12)))[0] == 0xd64e1e3efc5b8e9e
    assert ps.permute_g([P - 1] * 12)[11] == 0xe28e96f1ae5e60d3
    assert ps.permute_g(PLONKY2_TV12_IN) == PLONKY2_TV12_OUT
    assert ps.permute_b_fr([0, 1, 2, 3, 4])[0] == 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465   # circomlib
    assert ps.permute_b(list(range(12)))[0] == 0xd983775ce161c4e4
    rng = np.random.default_rng(2)
    st = rng.integers(0, P, size=(300, 12), dtype=np.uint64)
    st[0], st[1], st[2] = 0, P - 1, [P - 1, 0, 1, P - 2] * 3
    batch = ps.permute_g_batch(st)
    assert (batch == orc.poseidon_batch(st)).all()                      # oracle runs the reference's FAST form, pyref the naive one
    for i in range(0, 300, 37):
        assert [int(v) for v in batch[i]] == ps.permute_g([int(v) for v in st[i]])
    for i in range(3):
        assert ps.permute_b([int(v) for v in st[i]]) == [int(v) for v in orc.poseidon_b(st[i])]
    for kind in (0, 1):
        for n in (1, 4, 5, 8, 9, 17):
            row = [int(v) for v in rng.integers(0, P, size=n, dtype=np.uint64)]
            assert ps.hash_no_pad(row, kind) == [int(v) for v in orc.hash_no_pad(row, kind)]
        assert ps.two_to_one(st[5][:4], st[5][4:8], kind) == [int(v) for v in orc.two_to_one(st[5][:4], st[5][4:8], kind)]


def test_fold_of_arity_2_is_the_reference_formula():
    """fri_chip.rs:212-224 literally, against the general interpolation"""
    rng = np.random.default_rng(3)
    for _ in range(20):
        x = rf(rng, 1)
        ev = [(rf(rng), rf(rng)) for _ in range(2)]
        beta = (rf(rng), rf(rng))
        for bit_ in (0, 1):
            assert pfri.fold(x, bit_, 1, ev, beta) == pfri.fold_arity2_reference(x, bit_, ev, beta)


def test_fold_is_the_coefficient_fold():
    """for any arity: folding the values of P on a coset == sum_t beta^t P_t(x^r), P(X) = sum_t X^t P_t(X^r)"""
    rng = np.random.default_rng(4)
    for ab in (1, 2, 3, 4):
        r = 1 << ab
        coeffs = [(rf(rng), rf(rng)) for _ in range(4 * r)]
        x = rf(rng, 1)
        g = gl.root_of_unity(ab)
        beta = (rf(rng), rf(rng))

        def ev_at(pt):
            acc = (0, 0)
            for c in reversed(coeffs):
                acc = gl.e_add(gl.e_scale(acc, pt), c)
            return acc
        within = int(rng.integers(0, r))
        cs = x * pow(gl.inv(g), gl.bitrev(within, ab), P) % P
        leaf = [ev_at(cs * pow(g, gl.bitrev(j, ab), P) % P) for j in range(r)]        # leaf order: entry j at cs * g^bitrev(j)
        assert leaf[within] == ev_at(x)
        want = (0, 0)
        y = pow(x, r, P)
        for t in reversed(range(r)):
            pt = (0, 0)
            for c in reversed(coeffs[t::r]):
                pt = gl.e_add(gl.e_scale(pt, y), c)
            want = gl.e_add(gl.e_mul(want, beta), pt)
        assert pfri.fold(x, within, ab, leaf, beta) == want


@pytest.mark.parametrize("name", gp.names())
def test_oracle_reproduces_the_python_verdicts(svb, orc, name):
    """every golden record (valid, and one corrupted word each): accept bit, first-failure code and query round of the C
    oracle == those of tests/pyref/fri.py; the layouts agree; the host transcript of the product and the oracle's
    reproduce the recorded challenge fields"""
    G = gp.load(svb, name)
    params, L = G["params"], svb.api.make_layout(G["params"])
    oshape = orc.shape_from(params.to_shape())
    assert G["records"].shape[1] == L.record_words == orc.layout(oshape).record_words
    for rec, (accept, code, query) in zip(G["records"], G["verdicts"]):
        ok, ocode, oq = orc.fri_verify(oshape, np.ascontiguousarray(rec))
        assert (int(bool(ok)), ocode, max(oq, 0)) == (accept, code, query)
    bm = orc.fri_verify_batch(oshape, np.ascontiguousarray(G["records"]), nthreads=2)
    assert [bit(bm, i) for i in range(len(G["verdicts"]))] == [v[0] for v in G["verdicts"]]
    nch = G["meta"]["num_zs"]
    for i, rec in enumerate(G["base"]):
        pih = svb.public_inputs_hash(G["public_inputs"][i])
        assert [int(v) for v in pih] == ps.hash_no_pad(G["public_inputs"][i])
        for fill in (lambda r: svb.fri_challenges(params, r, G["circuit_digests"][i], pih, nch),
                     lambda r: orc.fri_challenges(oshape, r, G["circuit_digests"][i], pih, nch)):
            again = rec.copy()
            again[L.off_alpha:L.header_words] = 0
            fill(again)
            assert (again == rec).all()


@pytest.mark.parametrize("name", [n for n in gp.names() if not n.startswith("shape_a")])
def test_wire_bytes_of_the_python_writer(svb, orc, name):
    """product unpacker (host twin of the device gather) and oracle reader on the Python writer's bytes; product packer ==
    Python writer; Python reader on the product packer's bytes"""
    G = gp.load(svb, name)
    params, common, L = G["params"], G["common"], svb.api.make_layout(G["params"])
    blobs = G["blobs"]
    assert blobs.shape[1] == svb.wire_proof_bytes(common)
    stripped = G["base"].copy()
    stripped[:, L.off_alpha:L.header_words] = 0
    pih = []
    for i in range(blobs.shape[0]):       # every golden proof has its own "circuit", hence its own verifier-key cap
        recs, ph, pis, mal = svb.wire_unpack_batch(common, G["vk_caps"][i], np.ascontiguousarray(blobs[i]), nthreads=1)
        assert (recs[0] == stripped[i]).all() and not any(mal) and (pis[0] == G["public_inputs"][i]).all()
        pih.append(ph[0])
    vk = G["vk_caps"][0]
    assert (svb.wire_pack(common, G["base"], G["public_inputs"]) == blobs).all()
    pp = ppf.FriParams(params.degree_bits, params.config.rate_bits, params.config.cap_height, params.config.proof_of_work_bits,
                       params.config.num_query_rounds, list(params.reduction_arity_bits), list(params.oracle_num_polys), params.num_zs,
                       hiding=params.hiding, hash_kind=params.hash_kind)
    pc = ppf.Common(common.num_constants, common.num_routed_wires, common.num_wires, common.num_challenges,
                    common.num_partial_products, common.quotient_degree_factor, common.num_public_inputs)
    oshape = orc.shape_from(params.to_shape())
    for i in range(blobs.shape[0]):
        proof = ppf.read_proof(bytes(blobs[i]), pp, pc)
        assert ppf.write_proof(proof) == bytes(blobs[i]) and proof.public_inputs == [int(v) for v in G["public_inputs"][i]]
        rc, rec_o, pis_o, pih_o = orc.wire_read_proof(oshape, orc.common_from(common.to_c()), G["vk_caps"][i], blobs[i])
        assert rc == 0 and (rec_o == stripped[i]).all() and (pih_o == pih[i]).all()
    bad = bytearray(bytes(blobs[0]))
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    bad[q0 + 8 * L.leaf_len[0]] ^= 1                                      # the first Merkle-proof length byte
    with pytest.raises(ppf.Malformed):
        ppf.read_proof(bytes(bad), pp, pc)
    _, _, _, mal = svb.wire_unpack_batch(common, vk, np.frombuffer(bytes(bad), dtype=np.uint8).copy(), nthreads=1)
    assert list(mal) == [1]


@pytest.mark.parametrize("arity_bits,deg,hiding,kind", [(1, 7, False, 0), (2, 7, True, 0), (3, 8, False, 0), (4, 9, False, 0), (3, 6, False, 1)])
def test_python_verifier_on_the_products_prover(svb, orc, arity_bits, deg, hiding, kind):
    """live, both directions: the product's synthetic prover (csrc/host_side.cpp, folding with the SAME fri_fold the kernel
    runs) judged by the Python verifier and by the oracle -- equal verdicts, first-failure codes included, on valid and
    seeded-corrupt records; and a fresh Python-prover proof accepted by the oracle"""
    params = svb.api._params(deg, 3 if kind == 0 else 1, 2, 4, 5 if kind == 0 else 2, hiding=hiding, arity_bits=arity_bits,
                             oracle_num_polys=(9, 11, 6, 4), num_zs=2, hash_kind=kind)
    assert set(params.reduction_arity_bits) == {arity_bits}
    L = svb.api.make_layout(params)
    import full_prover as fp
    pp = fp.pyref_params(params)
    common = ppf.Common.for_widths(pp.oracle_num_polys, pp.num_zs, 0)
    oshape = orc.shape_from(params.to_shape())
    n = 10 if kind == 0 else 3
    recs = svb.synth_proofs(params, n, seed=0x601D + arity_bits, n_circuits=2)
    kinds = corrupt(recs, L, np.random.default_rng(9), every=2, num_steps=len(params.reduction_arity_bits))
    for i in range(n):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert pfri.verify_record(pp, common, recs[i]) == (bool(ok), code, max(q, 0)), (i, kinds.get(i))
        assert bool(ok) == (i not in kinds)
    if kind == 0:
        proof, vk_cap, cd, pr = ppr.prove_random(pp, common, seed=11)
        rec = np.array(ppf.to_record(pp, proof, vk_cap, pr.challenges, pch.zeta_next(pr.challenges["plonk_zeta"], pp.degree_bits)),
                       dtype=np.uint64)
        assert orc.fri_verify(oshape, rec)[0]
