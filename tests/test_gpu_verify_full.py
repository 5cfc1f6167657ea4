"""The complete verifier on the device (sv_verify_proofs_full): serialised COMPLETE proofs of toy circuits -- the
reference's recursion gate set included -- must be accepted, every corruption rejected with the right first-failure
class, and the verdicts must equal the CPU side's (oracle FRI verifier AND plonk identity)."""
import os

import numpy as np
import pytest

from common import bit
from test_full_proof import ROOT, build, cpu_verdicts
from test_plonk_check import CONFIGS, c_gates

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,hash_kind", [("one_selector", 0), ("two_selectors", 1), ("recursion_gate_set", 0)])
def test_full_verifier_matches_cpu_side(svb, orc, ctx, name, hash_kind):
    base = 3 if name != "recursion_gate_set" else 2
    B = build(svb, orc, name, base, seed=21, hash_kind=hash_kind)
    L, common = B["L"], B["common"]
    n = 41
    blob = np.stack([B["blob"][i % base] for i in range(n)])
    caps_end = 3 * 32 * L.ncap
    open_end = caps_end + 16 * (L.n0 + L.n1)
    blob[5, caps_end + 40] ^= 1                            # an opening            -> plonk identity (and FRI)
    blob[9, open_end + 8 * L.leaf_len[0] + 1 + 9] ^= 2      # a sibling             -> FRI only
    blob[17, open_end + 8 * L.leaf_len[0]] ^= 1             # a Merkle length byte  -> malformed
    blob[33, -3] ^= 1                                       # a public input        -> other challenges
    blob[40, 7] ^= 8                                        # the wires cap         -> other challenges
    fri, pl, opl, mal, _ = cpu_verdicts(svb, orc, B, blob)
    assert pl == opl
    want = [int(f and p and not m) for f, p, m in zip(fri, pl, mal)]
    assert [i for i in range(n) if not want[i]] == [5, 9, 17, 33, 40]
    bm, ff = ctx.verify_proofs_full(common, B["circuit"], B["vk_cap"], B["cd"], blob.reshape(-1), want_fail=True)
    assert [bit(bm, i) for i in range(n)] == want
    assert ff[17] == svb.FAIL_MALFORMED and ff[5] == svb.FAIL_PLONK and (ff[9] & 0xFF) == 3 and all(ff[i] == 0 for i in range(n) if want[i])
    assert pl[9] == 1 and fri[9] == 0
    # without the plonk identity the FRI-only entry point gives the FRI verdicts
    bm_fri = ctx.verify_proofs_wire(common, B["vk_cap"], B["cd"], blob.reshape(-1))
    assert [bit(bm_fri, i) for i in range(n)] == [int(f and not m) for f, m in zip(fri, mal)]
    # idempotent
    assert (ctx.verify_proofs_full(common, B["circuit"], B["vk_cap"], B["cd"], blob.reshape(-1)) == bm).all()


def test_full_verifier_with_fri_reduction_steps(svb, orc, ctx):
    B = build(svb, orc, "two_selectors", 2, seed=31, degree_bits=6)
    L = B["L"]
    blob = np.concatenate([B["blob"], B["blob"], B["blob"][:1]])
    blob[2, 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + 3] ^= 1      # the commit-phase cap
    blob[4, 3 * 32 * L.ncap + 5] ^= 1                           # an opening
    fri, pl, opl, mal, _ = cpu_verdicts(svb, orc, B, blob)
    want = [int(f and p and not m) for f, p, m in zip(fri, pl, mal)]
    assert want == [1, 1, 0, 1, 0]
    bm, ff = ctx.verify_proofs_full(B["common"], B["circuit"], B["vk_cap"], B["cd"], blob.reshape(-1), want_fail=True)
    assert [bit(bm, i) for i in range(5)] == want and ff[4] == svb.FAIL_PLONK and ff[2] not in (0, svb.FAIL_PLONK, svb.FAIL_MALFORMED)


def test_full_verifier_with_salted_leaves(svb, orc, ctx):
    B = build(svb, orc, "one_selector", 2, seed=41, hiding=True)
    L = B["L"]
    blob = np.concatenate([B["blob"], B["blob"][1:]])
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1)
    blob[2, q0 + 8 * L.leaf_len[0] + 1 + 32 * L.init_depth + 8 * (L.leaf_len[1] - 1)] ^= 1    # a salt limb of the wires leaf
    bm, ff = ctx.verify_proofs_full(B["common"], B["circuit"], B["vk_cap"], B["cd"], blob.reshape(-1), want_fail=True)
    assert [bit(bm, i) for i in range(3)] == [1, 1, 0] and (ff[2] & 0xFF) == 3


def test_plonk_verifier_chip_mirror(svb, orc, ctx):
    """The reference's three steps by name -- get_public_inputs_hash, get_challenges, verify_proof_with_challenges
    (chip/plonk/plonk_verifier_chip.rs) -- on records of complete proofs."""
    B = build(svb, orc, "all_gates", 3, seed=8)
    L = B["L"]
    chip = svb.PlonkVerifierChip(ctx, B["common"], B["circuit"])
    recs = B["recs"].copy()
    recs[:, L.off_alpha:L.header_words] = 0                       # the verifier derives every challenge itself
    recs[2, L.off_open0 + 9] ^= np.uint64(1)
    pihs = np.stack([chip.get_public_inputs_hash(B["pis"][i]) for i in range(3)])
    chals = np.stack([chip.get_challenges(pihs[i], B["cd"], recs[i]) for i in range(3)])
    assert (recs[:2] == B["recs"][:2]).all()
    assert chip.verify_proof_with_challenges(recs, pihs, chals) == [True, True, False]


def test_verify_batch_drop_in(svb, orc, ctx):
    """svb.verify_batch: byte strings + the reference's gate ids in, one bool per proof out."""
    B = build(svb, orc, "recursion_gate_set", 2, seed=5)
    C = B["C"]
    gate_ids = ["NoopGate", "ConstantGate { num_consts: 2 }", "PublicInputGate", "ArithmeticGate { num_ops: 20 }",
                "ArithmeticExtensionGate { num_ops: 10 }", "MulExtensionGate { num_ops: 13 }", "BaseSumGate { num_limbs: 63 } + Base: 2",
                "ReducingGate { num_coeffs: 43 }", "ReducingExtensionGate { num_coeffs: 32 }",
                "RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
                "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",
                "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>"]
    proofs = [B["blob"][0].tobytes(), B["blob"][1].tobytes(), bytes(bytearray(B["blob"][0].tobytes()[:-1]) + bytearray([B["blob"][0][-1] ^ 1]))]
    ok = svb.verify_batch(ctx, proofs, B["vk_cap"], B["cd"], B["common"], gate_ids, C.groups, C.k_is, C.num_gate_constraints)
    assert ok == [True, True, False]
    with pytest.raises(svb.SvError):
        svb.verify_batch(ctx, [proofs[0][:-8]], B["vk_cap"], B["cd"], B["common"], gate_ids, C.groups, C.k_is, C.num_gate_constraints)


def test_full_verifier_golden_blob(svb, ctx):
    import full_prover as fp
    g = np.load(os.path.join(ROOT, "tests", "golden", "full_proof_toy.npz"))
    C, params = fp.toy_setup(svb, CONFIGS[str(g["config"])])
    common = svb.CommonData.for_params(params, num_public_inputs=int(g["num_public_inputs"]), num_constants=C.num_constants)
    circuit = svb.make_plonk_circuit(common, c_gates(C), C.groups, C.k_is, C.num_gate_constraints)
    bm = ctx.verify_proofs_full(common, circuit, g["vk_cap"], g["circuit_digest"], g["blob"].reshape(-1))
    assert [bit(bm, i) for i in range(g["blob"].shape[0])] == [int(v) for v in g["accept"]]
