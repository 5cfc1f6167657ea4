"""The GPU hot path (fri_query_kernel through sv_fri_verify_batch / sv_fri_verify_batch_fs -- the entry points of
tests/test_gpu_parity.py) on COMPLETE proofs made by the independent pure-Python prover (tests/full_prover.py): a prover
that shares no code with the product's synthetic prover or with either verifier.  Records come from the committed wire
bytes through the CPU unpacker; the verdicts must equal the oracle's."""
import os

import numpy as np
import pytest

from common import bit
from test_full_proof import ROOT, build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,hash_kind,degree_bits,hiding", [("one_selector", 0, 4, False), ("two_selectors", 1, 4, False),
                                                            ("two_selectors", 0, 6, False), ("one_selector", 0, 4, True),
                                                            ("recursion_gate_set", 0, 4, False)])
def test_fri_kernel_accepts_python_prover_proofs(svb, orc, ctx, name, hash_kind, degree_bits, hiding):
    base = 2
    B = build(svb, orc, name, base, seed=77, hash_kind=hash_kind, degree_bits=degree_bits, hiding=hiding)
    params, L = B["params"], B["L"]
    n = 40
    recs = np.stack([B["recs"][i % base] for i in range(n)])
    rng = np.random.default_rng(1)
    bad = {}
    for i in range(3, n, 5):
        q = int(rng.integers(0, params.config.num_query_rounds))
        qb = L.header_words + q * L.query_words
        k = int(rng.integers(0, 4))
        if i % 2:
            recs[i, qb + L.q_off_init_sibs[k] + int(rng.integers(0, 4 * L.init_depth))] ^= np.uint64(1)
        else:
            recs[i, qb + L.q_off_init_evals[k] + int(rng.integers(0, L.leaf_len[k]))] ^= np.uint64(4)
        bad[i] = 1
    want = orc.fri_verify_batch(orc.shape_from(params.to_shape()), recs, nthreads=4)
    assert [i for i in range(n) if not bit(want, i)] == sorted(bad)
    got = ctx.fri_verify_batch(params, recs)
    assert (got == want).all()
    # the same through the device transcript: challenge fields stripped, derived on the GPU from hash(public inputs)
    stripped = recs.copy()
    stripped[:, L.off_alpha:L.header_words] = 0
    pih = np.stack([svb.public_inputs_hash(B["pis"][i % base]) for i in range(n)])
    got_fs = ctx.fri_verify_batch_fs(params, stripped, B["cd"], pih, num_challenges=B["C"].num_challenges)
    assert (got_fs == want).all()
