"""Plonk-level checks, CPU side: proofs from the independent pure-Python prover (tests/plonk_prover.py) must
be accepted by the product's check (plonk_check.hpp through sv_plonk_check_host -- the function the device kernel runs)
and by the oracle's statement-by-statement restatement (oracle/plonk.c); every corruption must be rejected by both; on
arbitrary (non-satisfying) inputs the two must agree."""
import numpy as np
import pytest

import plonk_prover as pp
from common import P, bit

CONFIGS = {
    "one_selector": dict(),
    "two_selectors": dict(groups=((0, 2), (2, 4))),
    "one_challenge_no_partials": dict(num_challenges=1, num_routed_wires=8, num_wires=8, quotient_degree_factor=8),
    "three_chunks": dict(num_routed_wires=16, num_wires=18, quotient_degree_factor=7, groups=((0, 3), (3, 4))),
    # every gate this library evaluates: + ArithmeticExtension, MulExtension, BaseSum, Reducing, ReducingExtension
    "all_gates": dict(num_routed_wires=24, num_wires=24, quotient_degree_factor=8, groups=((0, 5), (5, 9)),
                      extra_gates=((pp.GATE_ARITHMETIC_EXT, 3), (pp.GATE_MUL_EXT, 4), (pp.GATE_BASE_SUM, 5),
                                   (pp.GATE_REDUCING, 4), (pp.GATE_REDUCING_EXT, 3))),
    # the gate set and parameters of the reference's recursion circuits (gates/mod.rs:138-196) on the standard
    # recursion configuration: 135 wires, 80 routed, 2 challenges, quotient degree factor 8
    "recursion_gate_set": dict(num_routed_wires=80, num_wires=135, quotient_degree_factor=8,
                               groups=((0, 5), (5, 9), (9, 11), (11, 12)),
                               extra_gates=((pp.GATE_ARITHMETIC_EXT, 10), (pp.GATE_MUL_EXT, 13), (pp.GATE_BASE_SUM, 63),
                                            (pp.GATE_REDUCING, 43), (pp.GATE_REDUCING_EXT, 32),
                                            (pp.GATE_RANDOM_ACCESS, (4, 4, 2)), (pp.GATE_POSEIDON_MDS, 0), (pp.GATE_POSEIDON, 0))),
}


def c_gates(C):
    """the prover's gate list in the (kind, param, param2, param3) form of sv_plonk_gate"""
    return [(k,) + tuple(p) if isinstance(p, tuple) else (k, p) for k, p in C.gates]


def plonk_setup(svb, cfg):
    C = pp.Circuit(**cfg)
    widths = (C.num_constants + C.num_routed_wires, C.num_wires, C.num_challenges * (1 + C.num_partial_products),
              C.num_challenges * C.qdf)
    params = svb.api._params(C.degree_bits, 3, 1, 2, 2, oracle_num_polys=widths, num_zs=C.num_challenges)
    common = svb.CommonData.for_params(params, num_public_inputs=0, num_constants=C.num_constants)
    circuit = svb.make_plonk_circuit(common, c_gates(C), C.groups, C.k_is, C.num_gate_constraints)
    return C, params, circuit, svb.api.make_layout(params)


def to_records(L, proofs):
    recs = np.zeros((len(proofs), L.record_words), dtype=np.uint64)
    chal = []
    for i, pr in enumerate(proofs):
        o0 = np.array([w for e in pr["open0"] for w in e], dtype=np.uint64)
        o1 = np.array([w for e in pr["open1"] for w in e], dtype=np.uint64)
        assert o0.size == 2 * L.n0 and o1.size == 2 * L.n1
        recs[i, L.off_open0:L.off_open0 + o0.size] = o0
        recs[i, L.off_open1:L.off_open1 + o1.size] = o1
        recs[i, L.off_zeta:L.off_zeta + 2] = pr["zeta"]
        chal.append(pr["betas"] + pr["gammas"] + pr["alphas"])
    return recs, np.array(chal, dtype=np.uint64)


def oracle_bits(orc, ocirc, L, recs, pih, chal):
    return [orc.plonk_check(ocirc, r[L.off_open0:L.off_open0 + 2 * L.n0], r[L.off_open1:L.off_open1 + 2 * L.n1], pih[i], chal[i],
                            r[L.off_zeta:L.off_zeta + 2]) for i, r in enumerate(recs)]


@pytest.mark.parametrize("name", list(CONFIGS))
def test_prover_accepted_corruptions_rejected(svb, orc, name):
    C, params, circuit, L = plonk_setup(svb, CONFIGS[name])
    ocirc = orc.plonk_circuit_from(circuit)
    rng = np.random.default_rng(7)
    n = 3 if C.num_wires < 100 else 1
    pih = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    proofs = [pp.prove(C, 100 + i, [int(x) for x in pih[i]]) for i in range(n)]
    recs, chal = to_records(L, proofs)
    bm = svb.plonk_check_host(params, circuit, recs, pih, chal, nthreads=2)
    assert [bit(bm, i) for i in range(n)] == [1] * n
    assert oracle_bits(orc, ocirc, L, recs, pih, chal) == [1] * n
    # one flipped bit in every kind of input: constants, sigmas, wires, zs, partial products, quotient, zs_next, zeta,
    # beta, gamma, alpha, public-inputs hash
    nc, nr, nw, nch, npp, qdf = C.num_constants, C.num_routed_wires, C.num_wires, C.num_challenges, C.num_partial_products, C.qdf
    starts = np.cumsum([0, nc, nr, nw, nch, nch * npp, nch * qdf])
    spots = [("open", L.off_open0 + 2 * int(rng.integers(starts[k], starts[k + 1])) + int(rng.integers(0, 2)))
             for k in range(6) if starts[k + 1] > starts[k]]
    spots += [("open", L.off_open1 + int(rng.integers(0, 2 * nch))), ("open", L.off_zeta + 1)]
    spots += [("chal", k * nch + int(rng.integers(0, nch))) for k in range(3)] + [("pih", int(rng.integers(0, 4)))]
    bad_recs = np.repeat(recs[:1], len(spots), axis=0)
    bad_chal = np.repeat(chal[:1], len(spots), axis=0)
    bad_pih = np.repeat(pih[:1], len(spots), axis=0)
    for i, (what, at) in enumerate(spots):
        {"open": bad_recs, "chal": bad_chal, "pih": bad_pih}[what][i, at] ^= np.uint64(1 << int(rng.integers(0, 40)))
    bm = svb.plonk_check_host(params, circuit, bad_recs, bad_pih, bad_chal)
    got = [bit(bm, i) for i in range(len(spots))]
    want = oracle_bits(orc, ocirc, L, bad_recs, bad_pih, bad_chal)
    assert got == want
    # a wire that no gate constrains on any row's filter can still not change: every opening enters the identity
    assert got == [0] * len(spots), [s for s, g in zip(spots, got) if g]


def test_wrong_public_inputs_hash_is_rejected(svb, orc):
    C, params, circuit, L = plonk_setup(svb, CONFIGS["one_selector"])
    pih = np.array([[5, 6, 7, 8]], dtype=np.uint64)
    recs, chal = to_records(L, [pp.prove(C, 1, [5, 6, 7, 8])])
    assert bit(svb.plonk_check_host(params, circuit, recs, pih, chal), 0) == 1
    other = np.array([[5, 6, 7, 9]], dtype=np.uint64)
    assert bit(svb.plonk_check_host(params, circuit, recs, other, chal), 0) == 0


def test_product_and_oracle_agree_on_arbitrary_inputs(svb, orc):
    """Differential test in the style of the reference's own gate tests (gates/gate_test.rs:154-176): random inputs,
    the two restatements must give the same verdict -- including the corner cases below."""
    C, params, circuit, L = plonk_setup(svb, CONFIGS["two_selectors"])
    ocirc = orc.plonk_circuit_from(circuit)
    rng = np.random.default_rng(11)
    n = 40
    recs = np.zeros((n, L.record_words), dtype=np.uint64)
    recs[:, :L.header_words] = rng.integers(0, P, size=(n, L.header_words), dtype=np.uint64)
    pih = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    chal = rng.integers(0, P, size=(n, 3 * C.num_challenges), dtype=np.uint64)
    good, gchal = to_records(L, [pp.prove(C, 9, [int(x) for x in pih[0]])])
    recs[0], chal[0] = good[0], gchal[0]
    recs[1], chal[1] = good[0], gchal[0]
    recs[1, L.off_zeta:L.off_zeta + 2] = (1, 0)            # zeta = 1: L_0 denominator is zero -> reject (division by zero)
    recs[2], chal[2] = good[0], gchal[0]
    recs[2, L.off_open0 + 3] = np.uint64(P)                # non-canonical opening word
    recs[3], chal[3] = good[0], gchal[0]
    chal[3, 0] = np.uint64(P + 5)                          # non-canonical challenge
    bm = svb.plonk_check_host(params, circuit, recs, pih, chal, nthreads=3)
    got = [bit(bm, i) for i in range(n)]
    assert got == oracle_bits(orc, ocirc, L, recs, pih, chal)
    assert got[:4] == [1, 0, 0, 0] and sum(got) == 1


def test_unknown_gates_and_inconsistent_circuits_are_refused(svb):
    C, params, circuit, L = plonk_setup(svb, CONFIGS["one_selector"])
    common = svb.CommonData.for_params(params, num_public_inputs=0, num_constants=C.num_constants)
    with pytest.raises(svb.SvError):       # a gate kind this library does not evaluate (e.g. PoseidonGate): never a silent accept
        svb.make_plonk_circuit(common, C.gates + [(17, 0)], [(0, 5)], C.k_is, C.num_gate_constraints)
    with pytest.raises(svb.SvError):       # partial products inconsistent with the routed wires
        bad = svb.CommonData.for_params(params, num_public_inputs=0, num_constants=C.num_constants)
        bad.num_partial_products += 1
        svb.make_plonk_circuit(bad, C.gates, C.groups, C.k_is, C.num_gate_constraints)
    with pytest.raises(svb.SvError):       # too few constraint slots for the arithmetic gate
        svb.make_plonk_circuit(common, C.gates, C.groups, C.k_is, 2)
    # a gate parameter that would wrap the u32 wire / constraint products: refused, not "8 wires, 2 constraints"
    for gates in ([(svb.GATE_ARITHMETIC_EXT, 0x80000001)], [(svb.GATE_ARITHMETIC, 0x40000001)], [(svb.GATE_REDUCING, 0x55555556)],
                  [(svb.GATE_RANDOM_ACCESS, 4, 0x10000000, 0)], [(svb.GATE_BASE_SUM, 0xFFFFFFFF)], [(svb.GATE_MUL_EXT, 0x2AAAAAAB)]):
        with pytest.raises(svb.SvError):
            svb.make_plonk_circuit(common, C.gates + gates, [(0, len(C.gates) + 1)], C.k_is, C.num_gate_constraints)
    g = svb.plonk_gate_from_id("ArithmeticExtensionGate { num_ops: 2147483649 }")
    with pytest.raises(svb.SvError):
        svb.make_plonk_circuit(common, C.gates + [g], [(0, len(C.gates) + 1)], C.k_is, C.num_gate_constraints)
    # FRI shape and circuit must describe the same openings
    other = svb.api._params(C.degree_bits + 1, 3, 1, 2, 2, oracle_num_polys=tuple(params.oracle_num_polys), num_zs=C.num_challenges)
    with pytest.raises(svb.SvError):
        svb.plonk_check_host(other, circuit, np.zeros((1, svb.api.make_layout(other).record_words), dtype=np.uint64),
                             np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 3 * C.num_challenges), dtype=np.uint64))


@pytest.mark.parametrize("kind", [0, 1])
def test_plonk_challenges_match_oracle_and_the_fri_transcript(svb, orc, kind):
    """sv_plonk_challenges == the oracle's; and it is the prefix of the transcript that sv_fri_challenges runs: a proof
    whose wires cap changes gets other plonk challenges AND another zeta."""
    from common import tiny_params
    params = tiny_params(svb, hash_kind=kind, degree_bits=6, cap=1)
    recs = svb.synth_proofs(params, 2, seed=4, n_circuits=1)
    cds, pih = svb.synth_public_inputs(params, 2, seed=4, n_circuits=1)
    oshape = orc.shape_from(params.to_shape())
    L = svb.api.make_layout(params)
    for nc in (1, 2, 3):
        a = svb.plonk_challenges(params, recs[0], cds[0], pih[0], nc)
        b = orc.plonk_challenges(oshape, recs[0], cds[0], pih[0], nc)
        assert (a == b).all() and (a < P).all() and len(set(a.tolist())) == 3 * nc
    r = recs[1].copy()
    before = svb.plonk_challenges(params, r, cds[0], pih[1])
    r[L.off_init_caps + 4 * L.ncap] ^= np.uint64(1)        # first word of the wires cap
    after = svb.plonk_challenges(params, r, cds[0], pih[1])
    assert (before != after).all()


def test_gate_ids_of_the_reference(svb):
    """The id strings CustomGateRef::from matches (chip/plonk/gates/mod.rs:138-196), verbatim, map to the kinds and
    parameters the reference constructs; anything else is refused like its unimplemented!()."""
    ids = {
        "ArithmeticGate { num_ops: 20 }": (svb.GATE_ARITHMETIC, 20, 0, 0),
        "PublicInputGate": (svb.GATE_PUBLIC_INPUT, 0, 0, 0),
        "NoopGate": (svb.GATE_NOOP, 0, 0, 0),
        "ConstantGate { num_consts: 2 }": (svb.GATE_CONSTANT, 2, 0, 0),
        "BaseSumGate { num_limbs: 63 } + Base: 2": (svb.GATE_BASE_SUM, 63, 0, 0),
        "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>": (svb.GATE_POSEIDON, 0, 0, 0),
        "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>": (svb.GATE_POSEIDON_MDS, 0, 0, 0),
        "RandomAccessGate { bits: 1, num_copies: 20, num_extra_constants: 0, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>":
            (svb.GATE_RANDOM_ACCESS, 1, 20, 0),
        "RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>":
            (svb.GATE_RANDOM_ACCESS, 4, 4, 2),
        "ReducingExtensionGate { num_coeffs: 32 }": (svb.GATE_REDUCING_EXT, 32, 0, 0),
        "ReducingGate { num_coeffs: 43 }": (svb.GATE_REDUCING, 43, 0, 0),
        "ArithmeticExtensionGate { num_ops: 10 }": (svb.GATE_ARITHMETIC_EXT, 10, 0, 0),
        "MulExtensionGate { num_ops: 13 }": (svb.GATE_MUL_EXT, 13, 0, 0),
        "BaseSumGate { num_limbs: 4 } + Base: 2": (svb.GATE_BASE_SUM, 4, 0, 0),
    }
    for gid, want in ids.items():
        assert svb.plonk_gate_from_id(gid) == want
        assert svb.plonk_gate_from_id(gid + "  ") == want          # .trim_end()
    for bad in ("ExponentiationGate { num_power_bits: 66 }", "BaseSumGate { num_limbs: 63 } + Base: 4", "ArithmeticGate { num_ops: 20 } x",
                "CosetInterpolationGate", ""):
        with pytest.raises(svb.SvError):
            svb.plonk_gate_from_id(bad)
    # the ids above, in the reference's gate set, make a circuit the library accepts on the standard recursion configuration
    C, params, circuit, L = plonk_setup(svb, CONFIGS["recursion_gate_set"])
    from_ids = [svb.plonk_gate_from_id(g) for g in (
        "NoopGate", "ConstantGate { num_consts: 2 }", "PublicInputGate", "ArithmeticGate { num_ops: 20 }",
        "ArithmeticExtensionGate { num_ops: 10 }", "MulExtensionGate { num_ops: 13 }", "BaseSumGate { num_limbs: 63 } + Base: 2",
        "ReducingGate { num_coeffs: 43 }", "ReducingExtensionGate { num_coeffs: 32 }",
        "RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
        "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",
        "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>")]
    assert [tuple(g) for g in from_ids] == [tuple((list(g) + [0, 0])[:4]) for g in c_gates(C)]


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_more_witnesses_of_the_recursion_gate_set(svb, orc, seed):
    """Other witnesses (copy constraints, Poseidon swap bit, random-access indices, limb patterns differ per seed)."""
    C, params, circuit, L = plonk_setup(svb, CONFIGS["recursion_gate_set"])
    pih = np.random.default_rng(seed).integers(0, P, size=(1, 4), dtype=np.uint64)
    recs, chal = to_records(L, [pp.prove(C, 1000 + seed, [int(x) for x in pih[0]])])
    assert bit(svb.plonk_check_host(params, circuit, recs, pih, chal), 0) == 1
    assert oracle_bits(orc, orc.plonk_circuit_from(circuit), L, recs, pih, chal) == [1]
    # a gate the proof was not made for: swap two gate kinds of equal selector group -> the identity must break
    gates = c_gates(C)
    gates[5], gates[6] = gates[6], gates[5]
    common = svb.CommonData.for_params(params, num_public_inputs=0, num_constants=C.num_constants)
    other = svb.make_plonk_circuit(common, gates, C.groups, C.k_is, C.num_gate_constraints)
    assert bit(svb.plonk_check_host(params, other, recs, pih, chal), 0) == 0


@pytest.mark.parametrize("name", ["one_selector", "two_selectors", "recursion_gate_set"])
def test_circuit_from_common_data_is_the_hand_filled_circuit(svb, name):
    """sv_circuit_from_common_data (CommonData::from + CustomGateRef::from, types/common_data.rs:224-270, gates/mod.rs:138-196): from the
    gate id strings, SelectorsInfo and k_is of the prover's circuit -> byte for byte the sv_plonk_circuit / sv_fri_shape the tests
    fill by hand; an id outside the reference's table, a gate outside its selector group or a wrong k_is count is refused"""
    import ctypes
    C, params, circuit, L = plonk_setup(svb, CONFIGS[name])
    common = svb.CommonData.for_params(params, num_public_inputs=0, num_constants=C.num_constants)
    ids = [pp.gate_id(k, p) for k, p in C.gates]
    sel = [C.selector_index(i) for i in range(len(C.gates))]
    shape, got = svb.circuit_from_common_data(common, ids, sel, C.groups, C.k_is, C.num_gate_constraints)
    assert bytes(got) == bytes(circuit)
    assert bytes(shape) == bytes(params.to_shape())
    with pytest.raises(svb.SvError):
        svb.circuit_from_common_data(common, ids[:-1] + ["LookupGate { num_slots: 3 }"], sel, C.groups, C.k_is, C.num_gate_constraints)
    if len(C.groups) > 1:
        bad = list(sel)
        bad[0] = 1 - bad[0]
        with pytest.raises(svb.SvError):
            svb.circuit_from_common_data(common, ids, bad, C.groups, C.k_is, C.num_gate_constraints)
    with pytest.raises(svb.SvError):
        svb.circuit_from_common_data(common, ids, sel, C.groups, C.k_is[:-1], C.num_gate_constraints)
