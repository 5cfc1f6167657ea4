"""A COMPLETE proof of a toy circuit (tests/full_prover.py: commitments, Fiat-Shamir, openings, FRI opening proof -- the
reference's recursion gate set at the standard parameters included) through every stage of the verifier on the CPU
side: wire bytes -> unpack -> public-inputs hash -> transcript -> vanishing-polynomial identity (product host twin and
oracle) -> FRI query phase (oracle).  tests/test_gpu_verify_full.py runs the same proofs through sv_verify_proofs_full."""
import os

import numpy as np
import pytest

import full_prover as fp
from common import P, bit
from test_plonk_check import CONFIGS, c_gates

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(svb, orc, name, n, seed, n_pi=6, hash_kind=0, degree_bits=4, hiding=False):
    """n complete proofs of one circuit -> dict with everything the verifier side needs."""
    C, params = fp.toy_setup(svb, dict(CONFIGS[name], degree_bits=degree_bits), hiding=hiding)
    params.hash_kind = hash_kind
    rng = np.random.default_rng(seed)
    cd = rng.integers(0, P, size=4, dtype=np.uint64)
    L = svb.api.make_layout(params)
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi, num_constants=C.num_constants)
    circuit = svb.make_plonk_circuit(common, c_gates(C), C.groups, C.k_is, C.num_gate_constraints)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    # one circuit = one witness-independent set of constants / sigmas: the same seed for every proof would also repeat
    # the witness, so the proofs differ in their public inputs (hence transcripts) and share the verifier key
    outs = [fp.prove_full(C, params, seed, pis[i], cd) for i in range(n)]
    recs = np.stack([o[0] for o in outs])
    assert recs.shape[1] == L.record_words
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    assert (recs[:, L.off_init_caps:L.off_init_caps + 4 * L.ncap] == vk_cap).all()
    # the independent prover wrote its own wire bytes (tests/pyref/proof.py); the product's packer must produce the same
    blob = np.stack([np.frombuffer(o[1]["blob"], dtype=np.uint8) for o in outs])
    assert (svb.wire_pack(common, recs, pis) == blob).all()
    return dict(C=C, params=params, L=L, common=common, circuit=circuit, cd=cd, pis=pis, recs=recs, vk_cap=vk_cap,
                blob=blob.copy(), outs=[o[1] for o in outs])


def cpu_verdicts(svb, orc, B, blob):
    """The whole verifier on the CPU side of the boundary: (fri bits, plonk bits (product), plonk bits (oracle), malformed)."""
    params, L, C = B["params"], B["L"], B["C"]
    r2, pih, _, mal = svb.wire_unpack_batch(B["common"], B["vk_cap"], np.ascontiguousarray(blob).reshape(-1), nthreads=2)
    n = r2.shape[0]
    chal = np.stack([svb.plonk_challenges(params, r2[i], B["cd"], pih[i], C.num_challenges) for i in range(n)])
    for i in range(n):
        svb.fri_challenges(params, r2[i], B["cd"], pih[i], C.num_challenges)
    oshape = orc.shape_from(params.to_shape())
    fri = orc.fri_verify_batch(oshape, r2, nthreads=2)
    pl = svb.plonk_check_host(params, B["circuit"], r2, pih, chal, nthreads=2)
    ocirc = orc.plonk_circuit_from(B["circuit"])
    opl = [orc.plonk_check(ocirc, r[L.off_open0:L.off_open0 + 2 * L.n0], r[L.off_open1:L.off_open1 + 2 * L.n1], pih[i], chal[i],
                           r[L.off_zeta:L.off_zeta + 2]) for i, r in enumerate(r2)]
    return [bit(fri, i) for i in range(n)], [bit(pl, i) for i in range(n)], opl, list(mal), r2


@pytest.mark.parametrize("name,hash_kind", [("one_selector", 0), ("two_selectors", 1), ("recursion_gate_set", 0)])
def test_complete_proofs_verify_from_bytes(svb, orc, name, hash_kind):
    n = 3 if name != "recursion_gate_set" else 2
    B = build(svb, orc, name, n, seed=12, hash_kind=hash_kind)
    fri, pl, opl, mal, r2 = cpu_verdicts(svb, orc, B, B["blob"])
    assert fri == [1] * n and pl == [1] * n and opl == [1] * n and mal == [0] * n
    assert (r2 == B["recs"]).all()                      # bytes -> record -> transcript reproduces the prover's record


def test_complete_proof_with_fri_reduction_steps(svb, orc):
    """2^6 rows: the FRI instance has a commit phase (one arity-2 fold, fri_chip.rs:168-226,275-315) -- the third,
    independent implementation of the folding conventions next to the oracle's verifier and the product's prover."""
    B = build(svb, orc, "two_selectors", 2, seed=31, degree_bits=6)
    assert len(B["params"].reduction_arity_bits) == 1 and B["L"].step_depth[0] == B["L"].lde_bits - 1 - 1
    blob = np.concatenate([B["blob"], B["blob"][:1]])
    L = B["L"]
    step_caps_at = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1)
    blob[2, step_caps_at + 3] ^= 1                       # the commit-phase cap: other beta, and the step tree no longer matches
    fri, pl, opl, mal, r2 = cpu_verdicts(svb, orc, B, blob)
    assert fri == [1, 1, 0] and pl == [1, 1, 1] and opl == pl and mal == [0, 0, 0]
    assert (r2[:2] == B["recs"]).all()


def test_complete_proof_with_salted_leaves(svb, orc):
    """FriParams.hiding (the semaphore circuit's zero_knowledge configuration, plonky2_semaphore/access_set.rs:68-84):
    the wires / Z / quotient leaves carry 4 salt limbs that the Merkle proofs hash and the DEEP quotient ignores."""
    B = build(svb, orc, "one_selector", 2, seed=41, hiding=True)
    L = B["L"]
    assert list(L.leaf_len) == [B["common"].num_constants + 12, 13 + 4, 4 + 4, 16 + 4]
    fri, pl, opl, mal, r2 = cpu_verdicts(svb, orc, B, B["blob"])
    assert fri == [1, 1] and pl == [1, 1] and opl == [1, 1] and mal == [0, 0] and (r2 == B["recs"]).all()
    blob = B["blob"].copy()
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1)
    wires_salt = q0 + 8 * L.leaf_len[0] + 1 + 32 * L.init_depth + 8 * (L.leaf_len[1] - 1)     # last salt limb of the wires leaf, query 0
    blob[1, wires_salt] ^= 1
    fri, pl, _, _, _ = cpu_verdicts(svb, orc, B, blob)
    assert fri == [1, 0] and pl == [1, 1]               # the salt is hashed (Merkle proof fails) but is not an evaluation


def test_every_region_of_the_wire_bytes_is_checked(svb, orc):
    """One flipped bit anywhere in a valid proof must be caught by the stage that owns it; product and oracle agree."""
    B = build(svb, orc, "all_gates", 1, seed=3)
    L, params, common = B["L"], B["params"], B["common"]
    blob = B["blob"][0]
    nb = blob.size
    caps_end = 3 * 32 * L.ncap
    open_end = caps_end + 16 * (L.n0 + L.n1)
    q_bytes = (nb - open_end - 16 * params.final_poly_len() - 8 - 8 * common.num_public_inputs) // params.config.num_query_rounds
    fin = open_end + params.config.num_query_rounds * q_bytes
    rng = np.random.default_rng(5)
    # the Merkle-proof length bytes of a query round sit after each oracle's evaluations
    len_bytes, o = set(), 0
    for k in range(4):
        o += 8 * L.leaf_len[k]
        len_bytes.add(o)
        o += 1 + 32 * L.init_depth
    assert o == q_bytes

    def pick(lo, hi, avoid=()):
        while True:
            at = int(rng.integers(lo, hi))
            if at not in avoid:
                return at

    avoid = {open_end + q * q_bytes + b for q in range(params.config.num_query_rounds) for b in len_bytes}
    spots = {"wires_cap": pick(0, 32 * L.ncap), "zs_cap": pick(32 * L.ncap, 64 * L.ncap), "quotient_cap": pick(64 * L.ncap, caps_end),
             "opening": pick(caps_end, open_end), "query_round": pick(open_end, fin, avoid), "final_poly": pick(fin, fin + 16 * params.final_poly_len() - 16),
             "pow_witness": fin + 16 * params.final_poly_len(), "public_input": pick(nb - 8 * common.num_public_inputs, nb),
             "merkle_length_byte": open_end + min(len_bytes)}
    cases = np.repeat(blob[None, :], len(spots) + 1, axis=0)
    for i, at in enumerate(spots.values()):
        cases[i + 1, at] ^= np.uint8(1 << int(rng.integers(0, 7)))     # bit 7 of the top byte could make a word >= p: also fine
    fri, pl, opl, mal, _ = cpu_verdicts(svb, orc, B, cases)
    assert pl == opl
    assert (fri[0], pl[0], mal[0]) == (1, 1, 0)
    verdict = {name: bool(fri[i + 1] and pl[i + 1] and not mal[i + 1]) for i, name in enumerate(spots)}
    assert not any(verdict.values()), verdict
    names = list(spots)
    # who catches what: openings feed the identity; query data, final polynomial and PoW only the FRI proof; the length
    # byte is a parse error
    assert pl[1 + names.index("opening")] == 0
    assert pl[1 + names.index("query_round")] == 1 and fri[1 + names.index("query_round")] == 0
    assert pl[1 + names.index("final_poly")] == 1 and fri[1 + names.index("final_poly")] == 0
    assert mal[1 + names.index("merkle_length_byte")] == 1


def test_golden_full_proof(svb, orc):
    """tests/golden/full_proof_toy.npz (tools/gen_golden_full.py): committed wire bytes of complete proofs of the
    recursion-gate-set toy circuit; the CPU side of the verifier accepts them and rejects the committed corrupted copy."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "full_proof_toy.npz"))
    C, params = fp.toy_setup(svb, CONFIGS[str(g["config"])])
    common = svb.CommonData.for_params(params, num_public_inputs=int(g["num_public_inputs"]), num_constants=C.num_constants)
    circuit = svb.make_plonk_circuit(common, c_gates(C), C.groups, C.k_is, C.num_gate_constraints)
    B = dict(C=C, params=params, L=svb.api.make_layout(params), common=common, circuit=circuit, cd=g["circuit_digest"], vk_cap=g["vk_cap"])
    fri, pl, opl, mal, _ = cpu_verdicts(svb, orc, B, g["blob"])
    want = [int(v) for v in g["accept"]]
    assert [int(f and p and not m) for f, p, m in zip(fri, pl, mal)] == want and pl == opl
    assert want[0] == 1 and 0 in want
