"""oracle/fast (the AVX-512 arm of the CPU baseline, bench-only) against oracle.c bit for bit: the 8-lane permutation on corner
and random states, and the accept bitmaps on valid and seeded-corrupt proofs of several shapes (ragged query counts, salted
leaves, 2^k-ary reductions).  Skipped on a CPU without AVX-512 F + DQ."""
import numpy as np
import pytest

from common import P, corrupt


@pytest.fixture(scope="module")
def fast(orc):
    if not orc.fast_available():
        pytest.skip("no AVX-512 on this CPU")
    orc.fast_lib()
    return orc


def test_eight_lane_permutation(fast):
    rng = np.random.default_rng(3)
    for rep in range(40):
        st = rng.integers(0, P, size=(8, 12), dtype=np.uint64)
        if rep == 0:
            st[0], st[1], st[2], st[3] = 0, P - 1, [P - 1, 0, 1, P - 2] * 3, [0xFFFFFFFF, 0xFFFFFFFF00000000, 1 << 63] * 4
        assert (fast.fast_poseidon8(st) == fast.poseidon_batch(st)).all()
    assert int(fast.fast_poseidon8(np.zeros((8, 12), dtype=np.uint64))[5, 0]) == 0x3c18a9786cb0b359      # SURVEY 8c KAT


@pytest.mark.parametrize("kw", [dict(degree_bits=7, rate_bits=3, cap=2, queries=6), dict(degree_bits=6, rate_bits=2, cap=1, queries=9, hiding=True),
                                dict(degree_bits=8, rate_bits=3, cap=0, queries=17, arity_bits=3), dict(degree_bits=9, rate_bits=1, cap=3, queries=8, arity_bits=2),
                                dict(degree_bits=5, rate_bits=3, cap=1, queries=3)])
def test_bitmaps_equal_the_scalar_oracle(svb, fast, kw):
    params = svb.api._params(kw["degree_bits"], kw["rate_bits"], kw["cap"], 4, kw["queries"], hiding=kw.get("hiding", False),
                             arity_bits=kw.get("arity_bits", 1))
    L = svb.api.make_layout(params)
    recs = svb.synth_proofs(params, 40, seed=kw["degree_bits"], n_circuits=2)
    bad = corrupt(recs, L, np.random.default_rng(1), every=3, num_steps=len(params.reduction_arity_bits))
    osh = fast.shape_from(params.to_shape())
    want = fast.fri_verify_batch(osh, recs, nthreads=2)
    assert (fast.fast_fri_verify_batch(osh, recs, nthreads=3) == want).all()
    assert sum((int(want[i >> 5]) >> (i & 31)) & 1 for i in range(40)) == 40 - len(bad)


def test_hash_family_b_takes_the_scalar_path(svb, fast):
    params = svb.api._params(5, 1, 0, 2, 2, hash_kind=svb.HASH_POSEIDON_BN254)
    recs = svb.synth_proofs(params, 3, seed=1)
    recs[1, svb.api.make_layout(params).off_final_poly] ^= 1
    osh = fast.shape_from(params.to_shape())
    assert (fast.fast_fri_verify_batch(osh, recs) == fast.fri_verify_batch(osh, recs)).all()
