"""GPU parity for FRI reductions of arity 2^k: the fused kernel's fold (csrc/fri_fold.cuh: k successive
halvings) against the C oracle's (plonky2's barycentric compute_evaluation) on the product prover's proofs, valid and
corrupted, first-failure codes included; both hash families; the strategies the reference uses -- ConstantArityBits(1, 5)
(bn245_poseidon/plonky2_config.rs:84), ConstantArityBits(3, 5) (plonky2_semaphore/access_set.rs:124) -- and (2, 5), (4, 5);
the device transcript and the wire path with wider step leaves."""
import numpy as np
import pytest

from common import bit, corrupt

pytestmark = pytest.mark.gpu


def _check(svb, orc, ctx, params, n, seed):
    L = svb.api.make_layout(params)
    recs = svb.synth_proofs(params, n, seed=seed, n_circuits=2, nthreads=8)
    bad = corrupt(recs, L, np.random.default_rng(seed), every=4, num_steps=len(params.reduction_arity_bits))
    oshape = orc.shape_from(params.to_shape())
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    assert (bm == orc.fri_verify_batch(oshape, recs, nthreads=8)).all()
    for i in range(n):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert bit(bm, i) == int(bool(ok)) == (0 if i in bad else 1), (i, bad.get(i))
        assert int(ff[i]) == (0 if ok else (max(q, 0) << 8) | code), (i, bad.get(i), hex(int(ff[i])), code, q)
    return recs


@pytest.mark.parametrize("arity_bits", [1, 2, 3, 4])
@pytest.mark.parametrize("degree_bits,hiding,cap", [(8, False, 2), (10, True, 0), (7, False, 4)])
def test_fri_arity_small_shapes(svb, orc, ctx, arity_bits, degree_bits, hiding, cap):
    params = svb.api._params(degree_bits, 3, cap, 4, 7, hiding=hiding, arity_bits=arity_bits)
    assert params.reduction_arity_bits and set(params.reduction_arity_bits) == {arity_bits}
    _check(svb, orc, ctx, params, 48, seed=100 * arity_bits + degree_bits)


def test_fri_mixed_arities(svb, orc, ctx):
    params = svb.api._params(10, 2, 1, 3, 6, hiding=True)
    params.reduction_arity_bits = [2, 1, 4, 1]                # any list is a valid FriParams.reduction_arity_bits
    assert params.final_poly_len() == 4
    _check(svb, orc, ctx, params, 40, seed=5)


@pytest.mark.parametrize("arity_bits", [2, 3])
def test_fri_arity_hash_b(svb, orc, ctx, arity_bits):
    params = svb.api._params(7, 2, 1, 2, 4, arity_bits=arity_bits, hash_kind=svb.HASH_POSEIDON_BN254)
    _check(svb, orc, ctx, params, 12, seed=arity_bits)


def test_shape_a_with_the_semaphore_demo_strategy(svb, orc, ctx):
    """2^12 trace, blowup 8, 28 queries, cap 4, ConstantArityBits(3, 5): three 8-ary folds, final polynomial of 8"""
    params = svb.api._params(12, 3, 4, 16, 28, arity_bits=3)
    assert params.reduction_arity_bits == [3, 3, 3] and params.final_poly_len() == 8
    L = svb.api.make_layout(params)
    assert L.perms_per_query == 33 + 4 * 11 + (2 + 8) + (2 + 5) + (2 + 2)
    recs = _check(svb, orc, ctx, params, 64, seed=0xA8)
    # the same proofs through the device transcript
    cd, ph = svb.synth_public_inputs(params, 64, seed=0xA8, n_circuits=2)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=8)
    for c in range(2):
        sub = np.ascontiguousarray(recs[c::2]).copy()
        stripped = sub.copy()
        stripped[:, L.off_alpha:L.header_words] = 0
        filled = ctx.fri_challenges_batch(params, stripped.copy(), cd[c], np.ascontiguousarray(ph[c::2]))
        good = [i for i in range(sub.shape[0]) if bit(want, 2 * i + c)]
        assert (filled[good, L.off_alpha:L.header_words] == sub[good, L.off_alpha:L.header_words]).all()


def test_wire_path_with_arity_8(svb, orc, ctx):
    """plonky2 wire bytes with 8-ary query steps (16-word step leaves) through sv_verify_proofs_wire"""
    params = svb.api._params(9, 3, 2, 4, 6, arity_bits=3)
    common = svb.CommonData.for_params(params, num_public_inputs=3)
    L = svb.api.make_layout(params)
    n = 20
    rng = np.random.default_rng(8)
    pis = rng.integers(0, 0xFFFFFFFF00000001, size=(n, 3), dtype=np.uint64)
    pih = np.stack([svb.public_inputs_hash(pis[i]) for i in range(n)])
    recs = svb.synth_proofs(params, n, seed=31, n_circuits=1, pi_hashes=pih)
    cds, _ = svb.synth_public_inputs(params, n, seed=31, n_circuits=1)
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    blob = svb.wire_pack(common, recs, pis)
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    step0 = q0 + sum(8 * L.leaf_len[k] + 1 + 32 * L.init_depth for k in range(4))
    blob[3, step0 + 16 * 5] ^= 1                 # entry 5 of the first 8-ary step of query round 0
    blob[7, q0 + 9] ^= 1                         # an evaluation of oracle 0
    bm, ff = ctx.verify_proofs_wire(common, vk_cap, cds[0], blob, want_fail=True)
    assert [bit(bm, i) for i in range(n)] == [0 if i in (3, 7) else 1 for i in range(n)]
    r2, ph2, _, mal = svb.wire_unpack_batch(common, vk_cap, np.ascontiguousarray(blob).reshape(-1), nthreads=2)
    for i in range(n):
        svb.fri_challenges(params, r2[i], cds[0], ph2[i], common.num_challenges)
    oshape = orc.shape_from(params.to_shape())
    for i in (3, 7):
        ok, code, q = orc.fri_verify(oshape, r2[i])
        assert not ok and int(ff[i]) == (max(q, 0) << 8) | code
