"""fri_leaf_kernel + fri_query_kernel(leaf_digests) == the fused fri_query_kernel == the oracle.

Every path whose Fiat-Shamir transcript runs on the device hashes the oracle leaves BEFORE the challenges are known
(fri_leaf_kernel, csrc/fri_kernels.cuh) and starts the oracle-tree chains from those digests.  SVB_LEAF_SPLIT=0 keeps
the fused kernel; both must give the oracle's accept bits AND first-failure codes, including for a corrupted and for a
non-canonical leaf evaluation (the case the digest sentinel exists for)."""
import os

import numpy as np
import pytest

from common import P, bit, corrupt, tiny_params

pytestmark = pytest.mark.gpu


def _clear(recs, L, params):
    r = recs.copy()
    ns = len(params.reduction_arity_bits)
    for off, n in ((L.off_alpha, 2), (L.off_betas, 2 * ns), (L.off_pow_response, 1),
                   (L.off_indices, params.config.num_query_rounds), (L.off_zeta, 2), (L.off_zeta_next, 2)):
        r[:, off:off + n] = 0
    return r


def _corrupt_query_data(recs, L, params, rng):
    """Corruptions the transcript does not observe (so the derived challenges stay the proof's own)."""
    bad = {}
    nq = params.config.num_query_rounds
    for k, i in enumerate(range(2, recs.shape[0], 5)):
        q = int(rng.integers(0, nq))
        qb = L.header_words + q * L.query_words
        o = int(rng.integers(0, 4))
        kind = k % 5
        if kind == 0:      # leaf evaluation flipped
            recs[i, qb + L.q_off_init_evals[o] + int(rng.integers(0, L.leaf_len[o]))] ^= np.uint64(1 << int(rng.integers(0, 40)))
        elif kind == 1:    # leaf evaluation >= p
            recs[i, qb + L.q_off_init_evals[o] + int(rng.integers(0, L.leaf_len[o]))] = np.uint64(P + int(rng.integers(0, 99)))
        elif kind == 2:    # oracle sibling
            recs[i, qb + L.q_off_init_sibs[o] + int(rng.integers(0, 4 * L.init_depth))] ^= np.uint64(8)
        elif kind == 3:    # step evaluation
            recs[i, qb + L.q_off_step_evals[0] + int(rng.integers(0, 4))] ^= np.uint64(2)
        else:              # two queries of the same proof, the later one in an earlier-checked tree
            recs[i, qb + L.q_off_init_evals[3] + 1] ^= np.uint64(1)
            q2 = (q + 1) % nq
            recs[i, L.header_words + q2 * L.query_words + L.q_off_init_evals[0]] ^= np.uint64(1)
        bad[i] = kind
    return bad


@pytest.mark.parametrize("hiding,cap,degree_bits,kind", [(False, 3, 8, 0), (True, 1, 7, 0), (False, 2, 6, 1)])
def test_leaf_split_matches_fused_and_oracle(svb, orc, ctx, hiding, cap, degree_bits, kind):
    import torch
    params = tiny_params(svb, hiding=hiding, cap=cap, degree_bits=degree_bits, hash_kind=kind)
    L = svb.api.make_layout(params)
    base = 64 if kind else 128
    recs = svb.synth_proofs(params, base, seed=77 + cap, n_circuits=1)
    cd, ph = svb.synth_public_inputs(params, base, seed=77 + cap, n_circuits=1)
    bad = _corrupt_query_data(recs, L, params, np.random.default_rng(5))
    oshape = orc.shape_from(params.to_shape())
    want_bm = orc.fri_verify_batch(oshape, recs, nthreads=4)
    want_ff = np.zeros(base, dtype=np.uint32)
    for i in range(base):
        ok, code, q = orc.fri_verify(oshape, recs[i])     # the oracle's first failure: (query, code)
        assert bit(want_bm, i) == int(ok) == (0 if i in bad else 1), (i, bad.get(i))
        want_ff[i] = 0 if ok else ((max(q, 0) << 8) | code)
    stripped = _clear(recs, L, params)
    got = {}
    old = os.environ.get("SVB_LEAF_SPLIT")
    try:
        for mode in ("1", "0"):
            os.environ["SVB_LEAF_SPLIT"] = mode
            bm, ff = ctx.fri_verify_batch_fs(params, stripped, cd[0], ph, want_fail=True)
            assert (bm == want_bm).all(), mode
            assert (ff == want_ff).all(), (mode, np.nonzero(ff != want_ff)[0][:8])
            got[mode] = (bm.copy(), ff.copy())
            # device-resident batch of >= 4 096 proofs: the two-part transcript with the leaf kernel beside it
            reps = 4096 // base + 1
            n = reps * base
            d = torch.from_numpy(np.tile(stripped, (reps, 1)).view(np.int64)).cuda()
            dph = torch.from_numpy(np.tile(ph, (reps, 1)).view(np.int64)).cuda()
            dbm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
            dff = torch.zeros(n, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            ctx.fri_verify_batch_fs(params, d.data_ptr(), cd[0], dph.data_ptr(), n_proofs=n, accept_bitmap=dbm.data_ptr(),
                                    first_fail=dff.data_ptr(), mem=svb.MEM_DEVICE)
            ctx.synchronize()
            hbm = dbm.cpu().numpy().view(np.uint32)
            hff = dff.cpu().numpy().view(np.uint32)
            for i in range(n):
                assert bit(hbm, i) == bit(want_bm, i % base), (mode, i)
            assert (hff == np.tile(want_ff, reps)).all(), mode
    finally:
        if old is None:
            os.environ.pop("SVB_LEAF_SPLIT", None)
        else:
            os.environ["SVB_LEAF_SPLIT"] = old
    assert (got["1"][0] == got["0"][0]).all() and (got["1"][1] == got["0"][1]).all()


@pytest.mark.parametrize("chunk_mb,n", [(6, 560), (2, 645), (1, 200)])
def test_record_path_device_transcript_chunk_schedule(svb, orc, ctx, chunk_mb, n):
    """sv_fri_verify_batch_fs(SV_MEM_HOST) over several chunks: full chunks, the ramp-down at the end of the schedule and a
    chunk that straddles the two transcript parts -- verdicts and first-failure codes equal the oracle's."""
    params = tiny_params(svb, cap=2, degree_bits=7)
    L = svb.api.make_layout(params)
    base = 64
    recs = svb.synth_proofs(params, base, seed=31, n_circuits=1)
    cd, ph = svb.synth_public_inputs(params, base, seed=31, n_circuits=1)
    _corrupt_query_data(recs, L, params, np.random.default_rng(8))
    oshape = orc.shape_from(params.to_shape())
    ff1 = np.zeros(base, dtype=np.uint32)
    for i in range(base):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        ff1[i] = 0 if ok else ((max(q, 0) << 8) | code)
    reps = (n + base - 1) // base
    big = np.ascontiguousarray(np.tile(_clear(recs, L, params), (reps, 1))[:n])
    bph = np.ascontiguousarray(np.tile(ph, (reps, 1))[:n])
    old = os.environ.get("SVB_CHUNK_MB")
    os.environ["SVB_CHUNK_MB"] = str(chunk_mb)
    try:
        bm, ff = ctx.fri_verify_batch_fs(params, big, cd[0], bph, want_fail=True)
    finally:
        if old is None:
            del os.environ["SVB_CHUNK_MB"]
        else:
            os.environ["SVB_CHUNK_MB"] = old
    want_ff = np.tile(ff1, reps)[:n]
    assert (ff == want_ff).all(), np.nonzero(ff != want_ff)[0][:8]
    for i in range(n):
        assert bit(bm, i) == int(want_ff[i] == 0), i
    assert n & 31 == 0 or int(bm[-1]) >> (n & 31) == 0
