"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest

from common import P, PLONKY2_TV12_IN, PLONKY2_TV12_OUT, bit, corrupt, tiny_params

pytestmark = pytest.mark.gpu
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_permute_kat_and_random(svb, orc, ctx):
    rng = np.random.default_rng(0xB200)
    edge = np.array([[0] * 12, list(range(12)), [P - 1] * 12, [P - 1, 0, 1, P - 2] * 3], dtype=np.uint64)
    rnd = rng.integers(0, P, size=(4099, 12), dtype=np.uint64)   # ragged: not a multiple of the block
    states = np.concatenate([edge, rnd])
    got = ctx.poseidon_permute_batch(states)
    want = orc.poseidon_batch(states)
    assert (got.reshape(-1, 12) == want).all()
    assert int(got.reshape(-1, 12)[0, 0]) == 0x3c18a9786cb0b359
    assert int(got.reshape(-1, 12)[1, 0]) == 0xd64e1e3efc5b8e9e
    assert int(got.reshape(-1, 12)[2, 0]) == 0xbe0085cfc57a8357
    tv = ctx.poseidon_permute_batch(np.array(PLONKY2_TV12_IN, dtype=np.uint64))
    assert [int(x) for x in tv] == PLONKY2_TV12_OUT      # upstream plonky2 test_vectors12, random state


def test_permute_cooperative_mapping(svb, orc, ctx):
    """The lane-cooperative permutation of the device-side transcript (poseidon_g_coop2.cuh: partial section linearised,
    16 lanes per state) against the oracle: known answers, states of extremal halves (carries and borrows of its
    accumulators and of the LOOSE subtraction), ragged counts that leave a warp half empty."""
    rng = np.random.default_rng(0xC002)
    halves = [0, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000]
    edge = [[0] * 12, list(range(12)), [P - 1] * 12, [P - 1, 0, 1, P - 2] * 3, [0xFFFFFFFF] * 12, [0xFFFFFFFF00000000] * 12]
    for _ in range(200):
        edge.append([min(P - 1, (int(rng.choice(halves)) << 32) | int(rng.choice(halves))) for _ in range(12)])
    rnd = rng.integers(0, P, size=(4099, 12), dtype=np.uint64)
    states = np.concatenate([np.array(edge, dtype=np.uint64), rnd])
    want = orc.poseidon_batch(states)
    got = ctx.poseidon_permute_batch_coop(states).reshape(-1, 12)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, [(int(i), [hex(int(x)) for x in states[i]]) for i in bad[:3]]
    assert int(got[1, 0]) == 0xd64e1e3efc5b8e9e
    for n in (1, 2, 3, 17):                                  # one group, a half-empty warp, ...
        g = ctx.poseidon_permute_batch_coop(states[:n]).reshape(-1, 12)
        assert (g == want[:n]).all()
    tv = ctx.poseidon_permute_batch_coop(np.array(PLONKY2_TV12_IN, dtype=np.uint64))
    assert [int(x) for x in tv] == PLONKY2_TV12_OUT


def test_field_corner_cases(svb, ctx):
    """a*b + c mod p on the device against Python big integers, on operands whose 128-bit products hit
    the rare limb patterns of the reduction (borrow with a zero middle limb, carry out of the
    EPS multiply-add, both at once) -- random operands reach those with probability ~2^-32."""
    sp = [0, 1, 2, 0xFFFFFFFF, 0x100000000, 0x100000001, 0xFFFFFFFF00000000, P - 1, P, P + 1, 2**64 - 1,
          0xFFFFFFFEFFFFFFFF, 1 << 48, 5 << 48, 1 << 63, 0x00000001FFFFFFFF, 0x0000FFFF0000FFFF, 3 << 62,
          0x7FFFFFFF80000001, 0xFFFFFFFF, 0xFFFFFFFE00000001, 0x8000000000000001]
    a, b, c = [], [], []
    for x in sp:
        for y in sp:
            for z in sp[::3]:
                a.append(x); b.append(y); c.append(z)
    rng = np.random.default_rng(7)
    # products with chosen low / high limbs: a = 2^k, b = pattern
    for k in (16, 31, 32, 33, 47, 48, 63):
        for pat in (0xFFFFFFFF, 0xFFFFFFFF00000000, 0x00000001FFFFFFFF, 0x8000000080000000, 2**64 - 1):
            a.append(1 << k); b.append(pat); c.append(int(rng.integers(0, 2**63)))
    ra = rng.integers(0, 2**64, size=20000, dtype=np.uint64); rb = rng.integers(0, 2**64, size=20000, dtype=np.uint64)
    rc = rng.integers(0, 2**64, size=20000, dtype=np.uint64)
    A = np.concatenate([np.array(a, dtype=np.uint64), ra]); B = np.concatenate([np.array(b, dtype=np.uint64), rb])
    C = np.concatenate([np.array(c, dtype=np.uint64), rc])
    got = ctx.goldilocks_mul_add_batch(A, B, C)
    want = np.array([(int(x) * int(y) + int(z)) % P for x, y, z in zip(A, B, C)], dtype=np.uint64)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, [(hex(int(A[i])), hex(int(B[i])), hex(int(C[i])), hex(int(got[i])), hex(int(want[i]))) for i in bad[:5]]


def test_permute_empty(svb, ctx):
    out = ctx.poseidon_permute_batch(np.zeros((0, 12), dtype=np.uint64))
    assert out.size == 0


def _make_paths(orc, rng, n, leaf_len, depth, cap_height):
    """Random paths; the cap is built so that path 0..n-1 verify (each path gets its own cap slot
    only when cap_height allows; otherwise only the paths whose root lands in the cap verify)."""
    lw = (leaf_len + 3) & ~3
    rec = np.zeros((n, lw + 4 * depth), dtype=np.uint64)
    rec[:, :leaf_len] = rng.integers(0, P, size=(n, leaf_len), dtype=np.uint64)
    rec[:, lw:] = rng.integers(0, P, size=(n, 4 * depth), dtype=np.uint64)
    ncap = 1 << cap_height
    idx = rng.integers(0, 1 << (depth + cap_height), size=n, dtype=np.uint64)
    caps = rng.integers(0, P, size=(ncap, 4), dtype=np.uint64)
    # make the first ncap paths valid by writing their roots into their cap slot
    for i in range(min(n, ncap)):
        idx[i] = (np.uint64(i) << np.uint64(depth)) | (idx[i] & np.uint64((1 << depth) - 1))
        st = rec[i, :leaf_len].copy() if leaf_len <= 4 else orc.hash_no_pad(rec[i, :leaf_len])
        if leaf_len < 4:
            st = np.concatenate([st, np.zeros(4 - leaf_len, dtype=np.uint64)])
        for l in range(depth):
            sib = rec[i, lw + 4 * l: lw + 4 * l + 4]
            st = orc.two_to_one(sib, st) if (int(idx[i]) >> l) & 1 else orc.two_to_one(st, sib)
        caps[i] = st
    return rec, idx, caps


@pytest.mark.parametrize("leaf_len,depth,cap_height", [(4, 20, 0), (1, 3, 1), (3, 5, 2), (5, 4, 2), (8, 1, 0),
                                                       (9, 0, 3), (135, 11, 4), (84, 7, 4), (20, 6, 1), (16, 2, 0)])
def test_merkle_batch(svb, orc, ctx, leaf_len, depth, cap_height):
    rng = np.random.default_rng(leaf_len * 1000 + depth)
    n = 300
    rec, idx, caps = _make_paths(orc, rng, n, leaf_len, depth, cap_height)
    lw = (leaf_len + 3) & ~3
    # oracle wants unpadded leaves: feed it a compacted copy
    compact = np.concatenate([rec[:, :leaf_len], rec[:, lw:]], axis=1)
    # non-canonical word in a path that would otherwise verify
    if n > 0 and (1 << cap_height) > 1:
        rec[1, 0] = np.uint64(P)
        compact[1, 0] = np.uint64(P)
    got = ctx.merkle_verify_batch(leaf_len, depth, cap_height, rec, idx, caps)
    want = orc.merkle_verify_batch(compact, leaf_len, depth, idx, caps, cap_height)
    assert (got == want).all()
    assert got[0] == 1                      # a valid path is accepted
    assert got[(1 << cap_height):].sum() == 0  # random paths are rejected


@pytest.mark.parametrize("hiding,cap,degree_bits,rate_bits", [(False, 2, 7, 3), (True, 0, 8, 2), (False, 4, 6, 3), (True, 3, 9, 1)])
def test_fri_small_shapes(svb, orc, ctx, hiding, cap, degree_bits, rate_bits):
    params = tiny_params(svb, hiding=hiding, cap=cap, degree_bits=degree_bits, rate_bits=rate_bits)
    L = svb.api.make_layout(params)
    n = 77   # ragged: not a multiple of 32
    recs = svb.synth_proofs(params, n, seed=degree_bits * 31 + cap, n_circuits=2)
    rng = np.random.default_rng(5)
    bad = corrupt(recs, L, rng, every=4)
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=4)
    assert (bm == want).all()
    for i in range(n):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert bit(bm, i) == int(ok)
        assert bit(bm, i) == (0 if i in bad else 1), (i, bad.get(i), code)
        exp = 0 if ok else ((max(q, 0) << 8) | code)
        assert int(ff[i]) == exp, (i, bad.get(i), hex(int(ff[i])), hex(exp))


def test_fri_empty_and_single(svb, orc, ctx):
    params = tiny_params(svb)
    L = svb.api.make_layout(params)
    bm = ctx.fri_verify_batch(params, np.zeros((0, L.record_words), dtype=np.uint64), n_proofs=0)
    assert bm.size == 0
    recs = svb.synth_proofs(params, 1, seed=3)
    assert bit(ctx.fri_verify_batch(params, recs), 0) == 1


def test_fri_shape_a(svb, orc, ctx):
    """BASELINE configs[1] shape (2^12 trace, 28 queries, blowup 8, cap 4, PoW 16) on a batch the oracle
    finishes in seconds."""
    params = svb.SHAPE_A
    L = svb.api.make_layout(params)
    assert L.algo_bytes_per_query == 5240 and L.algo_bytes_shared == 10264 and L.perms_per_query == 126
    n = 96
    recs = svb.synth_proofs(params, n, seed=0xB2000002, n_circuits=2)
    bad = corrupt(recs, L, np.random.default_rng(11), every=8)
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=8)
    assert (bm == want).all()
    for i in range(n):
        assert bit(bm, i) == (0 if i in bad else 1), (i, bad.get(i))


def test_fri_semaphore_shape_salted(svb, orc, ctx):
    """BASELINE configs[0]: semaphore-shaped proof (zero_knowledge => hiding, salted leaves)."""
    params = svb.SHAPE_SEMAPHORE
    recs = svb.synth_proofs(params, 2, seed=1)
    bm = ctx.fri_verify_batch(params, recs)
    oshape = orc.shape_from(params.to_shape())
    assert (bm == orc.fri_verify_batch(oshape, recs, nthreads=2)).all()
    assert int(bm[0]) == 3


def test_fri_device_memory_path(svb, orc, ctx):
    """SV_MEM_DEVICE: records already resident, result left on the device (torch is only the allocator)."""
    import torch
    params = tiny_params(svb)
    L = svb.api.make_layout(params)
    n = 64
    recs = svb.synth_proofs(params, n, seed=9)
    bad = corrupt(recs, L, np.random.default_rng(2), every=8)
    d = torch.from_numpy(recs.view(np.int64)).cuda()
    bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    ctx.set_stream(stream.cuda_stream)
    ctx.fri_verify_batch(params, d.data_ptr(), n_proofs=n, accept_bitmap=bm.data_ptr(), mem=svb.MEM_DEVICE)
    ev = torch.cuda.Event(); ev.record(stream); ev.synchronize()   # the work really ran on `stream`
    ctx.set_stream(0)
    got = bm.cpu().numpy().view(np.uint32)
    want = orc.fri_verify_batch(orc.shape_from(params.to_shape()), recs, nthreads=2)
    assert (got == want).all()
    assert all(bit(got, i) == (0 if i in bad else 1) for i in range(n))


def test_full_size_tiled_batch(svb, orc, ctx):
    """configs[1] at full size (4 096 proofs): distinct base proofs tiled into physically distinct
    copies with a seeded 1/64 corruption pattern; the bitmap must equal the pattern, and the base
    proofs are checked against the oracle."""
    params = svb.SHAPE_A
    L = svb.api.make_layout(params)
    base = svb.synth_proofs(params, 16, seed=0xB2000002, n_circuits=2)
    oshape = orc.shape_from(params.to_shape())
    assert (orc.fri_verify_batch(oshape, base, nthreads=8) == np.array([0xFFFF], dtype=np.uint32)).all()
    n = 4096
    recs = np.ascontiguousarray(np.tile(base, (n // 16, 1)))
    bad = corrupt(recs, L, np.random.default_rng(64), every=64)
    bm = ctx.fri_verify_batch(params, recs)
    exp = np.zeros(n // 32, dtype=np.uint32)
    for i in range(n):
        if i not in bad:
            exp[i >> 5] |= np.uint32(1 << (i & 31))
    assert (bm == exp).all()
    # idempotence: the same call again gives the same bitmap
    assert (ctx.fri_verify_batch(params, recs) == exp).all()


def _clear_challenges(recs, L, params):
    """zero every field the transcript derives"""
    r = recs.copy()
    ns = len(params.reduction_arity_bits)
    for off, n in ((L.off_alpha, 2), (L.off_betas, 2 * ns), (L.off_pow_response, 1),
                   (L.off_indices, params.config.num_query_rounds), (L.off_zeta, 2), (L.off_zeta_next, 2)):
        r[:, off:off + n] = 0
    return r


@pytest.mark.parametrize("hiding,cap,degree_bits", [(False, 2, 7), (True, 0, 8), (False, 4, 6)])
def test_device_transcript_matches_host_and_oracle(svb, orc, ctx, hiding, cap, degree_bits):
    """sv_fri_challenges_batch (one GPU thread per proof) == host sv_fri_challenges == oracle transcript."""
    params = tiny_params(svb, hiding=hiding, cap=cap, degree_bits=degree_bits)
    L = svb.api.make_layout(params)
    n = 45
    recs = svb.synth_proofs(params, n, seed=21 + cap, n_circuits=1)
    rng = np.random.default_rng(99)
    cd = rng.integers(0, P, size=4, dtype=np.uint64)
    ph = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    dev = _clear_challenges(recs, L, params)
    ctx.fri_challenges_batch(params, dev, cd, ph)
    oshape = orc.shape_from(params.to_shape())
    for i in range(n):
        h = _clear_challenges(recs[i:i + 1], L, params)[0]
        svb.fri_challenges(params, h, cd, ph[i])
        o = _clear_challenges(recs[i:i + 1], L, params)[0]
        orc.fri_challenges(oshape, o, cd, ph[i])
        assert (h == o).all()
        assert (dev[i] == h).all(), i
    # and with the inputs the prover used, the transcript reproduces the proofs' own challenges
    cd0, ph0 = svb.synth_public_inputs(params, n, seed=21 + cap, n_circuits=1)
    dev = _clear_challenges(recs, L, params)
    ctx.fri_challenges_batch(params, dev, cd0[0], ph0)
    assert (dev == recs).all()


def test_verify_with_device_transcript(svb, orc, ctx):
    """sv_fri_verify_batch_fs: challenge fields are ignored on input and derived on the device; the accept
    bitmap equals the oracle's on the fully specified records (host and device memory paths)."""
    import torch
    params = tiny_params(svb, cap=3, degree_bits=8)
    L = svb.api.make_layout(params)
    n = 96
    recs = svb.synth_proofs(params, n, seed=5, n_circuits=1)
    cd, ph = svb.synth_public_inputs(params, n, seed=5, n_circuits=1)
    # corrupt data the transcript does not observe (siblings / leaf evals / step evals), so that the
    # derived challenges stay those of the original proof and the oracle can check the same records
    rng = np.random.default_rng(17)
    bad = {}
    for i in range(3, n, 7):
        q = int(rng.integers(0, params.config.num_query_rounds))
        qb = L.header_words + q * L.query_words
        which = i % 3
        if which == 0:
            recs[i, qb + L.q_off_init_sibs[1] + 3] ^= np.uint64(2)
        elif which == 1:
            recs[i, qb + L.q_off_init_evals[0] + 1] ^= np.uint64(1)
        else:
            recs[i, qb + L.q_off_step_evals[0] + 2] ^= np.uint64(4)
        bad[i] = which
    ph[10, 0] ^= np.uint64(1)            # wrong public-input hash: different challenges => reject
    bad[10] = "pi"
    want = orc.fri_verify_batch(orc.shape_from(params.to_shape()), recs, nthreads=4)
    stripped = _clear_challenges(recs, L, params)
    keep = stripped.copy()
    bm, ff = ctx.fri_verify_batch_fs(params, stripped, cd[0], ph, want_fail=True)
    assert (stripped == keep).all()                       # SV_MEM_HOST: host records untouched
    for i in range(n):
        exp = 0 if i in bad else 1
        assert bit(bm, i) == exp, (i, bad.get(i))
        if i != 10:
            assert bit(bm, i) == bit(want, i)
    # device-resident records
    d = torch.from_numpy(stripped.view(np.int64)).cuda()
    dph = torch.from_numpy(ph.view(np.int64)).cuda()
    dbm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.fri_verify_batch_fs(params, d.data_ptr(), cd[0], dph.data_ptr(), n_proofs=n, accept_bitmap=dbm.data_ptr(), mem=svb.MEM_DEVICE)
    ctx.synchronize()
    assert (dbm.cpu().numpy().view(np.uint32) == bm).all()
    # in device mode the challenge fields were written in place
    back = d.cpu().numpy().view(np.uint64)
    good = [i for i in range(n) if i != 10]
    assert (back[good] == recs[good]).all()


def test_config3_full_size_shape_b(svb, orc, ctx):
    """BASELINE configs[2] at full size: 65 536 proofs of shape B (2^20 trace, 84 queries, blowup 4) resident
    on one GPU (54 GB).  The base proof is checked against the oracle, tiled on the device into
    physically distinct records, 1/512 of them corrupted (a different class each); the accept bitmap
    must equal the pattern, and a second run must reproduce it (idempotence)."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    params = svb.SHAPE_B
    L = svb.api.make_layout(params)
    assert L.algo_bytes_per_query == 9624 and L.algo_bytes_shared == 14360 and L.perms_per_query == 255
    n = 65536
    rw = L.record_words
    if free < n * rw * 8 + (4 << 30):
        n = int((free - (4 << 30)) // (rw * 8)) // 1024 * 1024
        assert n >= 1024, "not enough device memory for a meaningful shape-B batch"
    # the base proof is a committed fixture (tools/gen_golden.py; the shape-B prover run takes minutes on the host)
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "fri_full_shapes.npz"))
    base = np.ascontiguousarray(fx["shape_b_record"]).reshape(1, rw)
    oshape = orc.shape_from(params.to_shape())
    assert int(orc.fri_verify_batch(oshape, base, nthreads=1)[0]) == 1
    d = torch.empty((n, rw), dtype=torch.int64, device="cuda")
    d[:1].copy_(torch.from_numpy(base.view(np.int64)))
    k = 1
    while k < n:
        c = min(k, n - k)
        d[k:k + c].copy_(d[:c])
        k += c
    rng = np.random.default_rng(3)
    rows, cols, exp = [], [], np.full(n // 32, 0xFFFFFFFF, dtype=np.uint32)
    ns = len(params.reduction_arity_bits)
    for j, i in enumerate(range(7, n, 512)):
        q = int(rng.integers(0, params.config.num_query_rounds))
        qb = L.header_words + q * L.query_words
        col = [qb + L.q_off_init_sibs[j % 4] + int(rng.integers(0, 4 * L.init_depth)),
               qb + L.q_off_init_evals[j % 4] + int(rng.integers(0, L.leaf_len[j % 4])),
               qb + L.q_off_step_evals[j % ns] + int(rng.integers(0, 4)),
               qb + L.q_off_step_sibs[j % ns] + int(rng.integers(0, 4 * L.step_depth[j % ns])),
               L.off_final_poly + int(rng.integers(0, 64))][j % 5]
        rows.append(i); cols.append(col)
        exp[i >> 5] &= np.uint32(~(1 << (i & 31)) & 0xFFFFFFFF)
    r_t, c_t = torch.tensor(rows, device="cuda"), torch.tensor(cols, device="cuda")
    d[r_t, c_t] = d[r_t, c_t] ^ 1
    bm = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):
        bm.zero_()
        torch.cuda.synchronize()
        ctx.fri_verify_batch(params, d.data_ptr(), n_proofs=n, accept_bitmap=bm.data_ptr(), mem=svb.MEM_DEVICE)
        ctx.synchronize()
        assert (bm.cpu().numpy().view(np.uint32) == exp).all()
    del d
    torch.cuda.empty_cache()


def test_config5_full_size_merkle_paths(svb, orc, ctx):
    """BASELINE configs[4]: 2^24 independent Merkle paths x depth 20, 4-limb leaves, cap_height 0.  The
    first 2^13 paths are made valid against a per-path expected root through the oracle... which needs
    one cap, so instead: all paths share one random cap (=> reject), a sampled subset is compared with
    the oracle, and the subset's own roots (computed by the oracle) accept when supplied as the cap."""
    import torch
    n, depth = 1 << 24, 20
    g = torch.Generator(device="cuda"); g.manual_seed(0xB2000005)
    paths = torch.randint(0, 1 << 62, (n, 4 + 4 * depth), dtype=torch.int64, device="cuda", generator=g)
    idx = torch.randint(0, 1 << depth, (n,), dtype=torch.int64, device="cuda", generator=g)
    # plant ONE known path: its root (from the oracle) becomes the cap, so exactly the copies of that path accept
    p0 = paths[0].cpu().numpy().view(np.uint64); i0 = int(idx[0])
    st = p0[:4].copy()
    for l in range(depth):
        sib = p0[4 + 4 * l: 8 + 4 * l]
        st = orc.two_to_one(sib, st) if (i0 >> l) & 1 else orc.two_to_one(st, sib)
    cap = torch.from_numpy(st.view(np.int64)).cuda()
    plant = torch.arange(0, n, 4099, device="cuda")
    paths[plant] = paths[0].clone()
    idx[plant] = idx[0].clone()
    bad = plant[1::2]
    paths[bad, 4 + 4 * 7 + 2] ^= 1          # one sibling limb flipped in every other planted copy
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.merkle_verify_batch(4, depth, 0, paths.data_ptr(), idx.data_ptr(), cap.data_ptr(), ok.data_ptr(), n=n, mem=svb.MEM_DEVICE)
    ctx.synchronize()
    want = torch.zeros(n, dtype=torch.uint8, device="cuda")
    want[plant[0::2]] = 1
    assert torch.equal(ok, want)
    # sampled subset against the oracle
    sub = 2048
    hp = paths[:sub].cpu().numpy().view(np.uint64); hi = idx[:sub].cpu().numpy().view(np.uint64)
    o = orc.merkle_verify_batch(hp, 4, depth, hi, st.reshape(1, 4), 0)
    assert (ok[:sub].cpu().numpy() == o).all()


# ---- hash family B: Poseidon over BN254 Fr wrapped around 12 Goldilocks limbs ----
def test_permute_b_golden_and_random(svb, orc, ctx):
    import json
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "poseidon_b.json")))
    ins = np.array([[int(x, 16) for x in c["in"]] for c in d["wrapped_permutation"]], dtype=np.uint64)
    outs = np.array([[int(x, 16) for x in c["out"]] for c in d["wrapped_permutation"]], dtype=np.uint64)
    got = ctx.poseidon_permute_batch(ins, hash_kind=svb.HASH_POSEIDON_BN254).reshape(-1, 12)
    assert (got == outs).all()
    assert int(got[1, 0]) == 0xd983775ce161c4e4 and int(got[1, 11]) == 0x6bf843b27c9d3fbb      # SURVEY 8c KAT
    rng = np.random.default_rng(0xB254)
    edge = np.array([[P - 1, 0, 1, P - 2] * 3, [P - 1] * 12, [0] * 11 + [P - 1]], dtype=np.uint64)
    rnd = rng.integers(0, P, size=(517, 12), dtype=np.uint64)      # ragged
    states = np.concatenate([edge, rnd])
    got = ctx.poseidon_permute_batch(states, hash_kind=svb.HASH_POSEIDON_BN254).reshape(-1, 12)
    want = np.array([orc.poseidon_b(s) for s in states], dtype=np.uint64)
    assert (got == want).all()
    assert (got < np.uint64(P)).all()


@pytest.mark.parametrize("leaf_len,depth,cap_height", [(4, 6, 0), (3, 2, 1), (9, 3, 2), (20, 4, 1), (135, 3, 0)])
def test_merkle_batch_hash_b(svb, orc, ctx, leaf_len, depth, cap_height):
    rng = np.random.default_rng(leaf_len * 77 + depth)
    n = 70
    lw = (leaf_len + 3) & ~3
    rec = np.zeros((n, lw + 4 * depth), dtype=np.uint64)
    rec[:, :leaf_len] = rng.integers(0, P, size=(n, leaf_len), dtype=np.uint64)
    rec[:, lw:] = rng.integers(0, P, size=(n, 4 * depth), dtype=np.uint64)
    ncap = 1 << cap_height
    idx = rng.integers(0, 1 << (depth + cap_height), size=n, dtype=np.uint64)
    caps = rng.integers(0, P, size=(ncap, 4), dtype=np.uint64)
    for i in range(min(n, ncap)):
        idx[i] = (np.uint64(i) << np.uint64(depth)) | (idx[i] & np.uint64((1 << depth) - 1))
        st = rec[i, :leaf_len].copy() if leaf_len <= 4 else orc.hash_no_pad(rec[i, :leaf_len], kind=1)
        if leaf_len < 4:
            st = np.concatenate([st, np.zeros(4 - leaf_len, dtype=np.uint64)])
        for l in range(depth):
            sib = rec[i, lw + 4 * l: lw + 4 * l + 4]
            st = orc.two_to_one(sib, st, kind=1) if (int(idx[i]) >> l) & 1 else orc.two_to_one(st, sib, kind=1)
        caps[i] = st
    compact = np.concatenate([rec[:, :leaf_len], rec[:, lw:]], axis=1)
    got = ctx.merkle_verify_batch(leaf_len, depth, cap_height, rec, idx, caps, hash_kind=svb.HASH_POSEIDON_BN254)
    want = orc.merkle_verify_batch(compact, leaf_len, depth, idx, caps, cap_height, kind=1)
    assert (got == want).all()
    assert got[0] == 1 and got[ncap:].sum() == 0
    # the same paths under the Goldilocks hash are rejected
    assert ctx.merkle_verify_batch(leaf_len, depth, cap_height, rec, idx, caps)[0] == 0 or depth == 0


@pytest.mark.parametrize("hiding,cap,degree_bits", [(False, 0, 6), (True, 2, 7)])
def test_fri_hash_b_small_shapes(svb, orc, ctx, hiding, cap, degree_bits):
    """The whole FRI query phase under hash family B (outer-proof configuration: cap_height 0)."""
    params = tiny_params(svb, hiding=hiding, cap=cap, degree_bits=degree_bits, hash_kind=svb.HASH_POSEIDON_BN254)
    L = svb.api.make_layout(params)
    n = 40
    recs = svb.synth_proofs(params, n, seed=degree_bits, n_circuits=1)
    bad = corrupt(recs, L, np.random.default_rng(8), every=4)
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=8)
    assert (bm == want).all()
    for i in range(n):
        assert bit(bm, i) == (0 if i in bad else 1), (i, bad.get(i))
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert int(ff[i]) == (0 if ok else ((max(q, 0) << 8) | code))
    # device-side transcript with the BN254 sponge
    cd, ph = svb.synth_public_inputs(params, n, seed=degree_bits, n_circuits=1)
    fresh = svb.synth_proofs(params, 8, seed=degree_bits, n_circuits=1)
    dev = _clear_challenges(fresh, L, params)
    ctx.fri_challenges_batch(params, dev, cd[0], ph[:8])
    assert (dev == fresh).all()


def test_error_convention(svb, ctx):
    """Bad arguments are errors (negative return + message); an invalid proof never is."""
    import ctypes
    lib = svb.lib()
    params = tiny_params(svb)
    L = svb.api.make_layout(params)
    recs = np.zeros((2, L.record_words), dtype=np.uint64)           # all-zero records: invalid proofs, not errors
    bm = ctx.fri_verify_batch(params, recs)
    assert int(bm[0]) == 0
    s = params.to_shape()
    s.hash_kind = 7
    out = np.zeros(1, dtype=np.uint32)
    rc = lib.sv_fri_verify_batch(ctx._h, ctypes.byref(s), 2, recs.ctypes.data, out.ctypes.data, None, svb.MEM_HOST)
    assert rc < 0 and b"hash_kind" in lib.sv_last_error(ctx._h)
    s = params.to_shape()
    s.num_steps = 40                                                 # > SV_MAX_STEPS
    rc = lib.sv_fri_verify_batch(ctx._h, ctypes.byref(s), 2, recs.ctypes.data, out.ctypes.data, None, svb.MEM_HOST)
    assert rc < 0
    s = params.to_shape()
    rc = lib.sv_fri_verify_batch(ctx._h, ctypes.byref(s), 2, recs.ctypes.data, out.ctypes.data, None, 5)   # unknown `mem`
    assert rc < 0 and b"SV_MEM" in lib.sv_last_error(ctx._h)
    with pytest.raises(svb.SvError):
        ctx.merkle_verify_batch(0, 3, 0, np.zeros((1, 16), dtype=np.uint64), np.zeros(1, dtype=np.uint64), np.zeros(4, dtype=np.uint64))
    with pytest.raises(svb.SvError):
        ctx.poseidon_permute_batch(np.zeros((1, 12), dtype=np.uint64), hash_kind=9)
    # the context is still usable after errors
    assert int(ctx.poseidon_permute_batch(np.arange(12, dtype=np.uint64))[0]) == 0xd64e1e3efc5b8e9e


@pytest.mark.parametrize("degree_bits,rate_bits,cap,queries", [(5, 3, 2, 4), (5, 1, 0, 3), (6, 1, 6, 5), (7, 1, 6, 2)])
def test_fri_edge_shapes(svb, orc, ctx, degree_bits, rate_bits, cap, queries):
    """Degenerate shapes: no reduction step at all (degree_bits = 5 with a 32-coefficient final polynomial),
    step trees of depth 0 (the coset pair IS the cap entry), caps as tall as the tree allows."""
    params = tiny_params(svb, cap=cap, queries=queries, degree_bits=degree_bits, rate_bits=rate_bits)
    L = svb.api.make_layout(params)
    ns = len(params.reduction_arity_bits)
    n = 37
    recs = svb.synth_proofs(params, n, seed=degree_bits * 100 + cap, n_circuits=2)
    bad = corrupt(recs, L, np.random.default_rng(4), every=4, num_steps=ns)
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=4)
    assert (bm == want).all()
    for i in range(n):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert bit(bm, i) == int(ok)
        assert int(ff[i]) == (0 if ok else ((max(q, 0) << 8) | code)), (i, bad.get(i))
    assert sum(bit(bm, i) for i in range(n)) >= n - len(bad)      # every untouched proof is accepted


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("leaf_len,log_n,cap_height", [(4, 5, 0), (3, 3, 1), (9, 4, 2), (135, 3, 0), (20, 2, 2)])
def test_merkle_tree_build_matches_oracle(svb, orc, ctx, kind, leaf_len, log_n, cap_height):
    """sv_merkle_tree_build (leaf digests + every level down to the cap) against the oracle's hash_no_pad /
    two_to_one, for both hash families."""
    rng = np.random.default_rng(leaf_len * 10 + log_n)
    n = 1 << log_n
    leaves = rng.integers(0, P, size=(n, leaf_len), dtype=np.uint64)
    layers = ctx.merkle_tree_build(leaves, leaf_len, cap_height, hash_kind=kind)
    assert len(layers) == log_n - cap_height + 1 and layers[-1].shape == (1 << cap_height, 4)
    want = []
    for i in range(n):
        if leaf_len <= 4:
            d = np.zeros(4, dtype=np.uint64); d[:leaf_len] = leaves[i]
        else:
            d = orc.hash_no_pad(leaves[i], kind=kind)
        want.append(d)
    want = np.array(want, dtype=np.uint64)
    assert (layers[0] == want).all()
    for lvl in range(1, len(layers)):
        want = np.array([orc.two_to_one(want[2 * j], want[2 * j + 1], kind=kind) for j in range(len(want) // 2)], dtype=np.uint64)
        assert (layers[lvl] == want).all()


def test_merkle_build_then_verify_roundtrip_large(svb, ctx):
    """Size-independent property at scale: build a 2^18-leaf tree on the device, open 2^16 random leaves (paths
    read out of the layers), verify them all on the device -> every path accepts; flip one bit in a quarter
    of them -> exactly those reject."""
    import torch
    log_n, leaf_len, cap_height = 18, 7, 3
    n = 1 << log_n
    g = torch.Generator(device="cuda"); g.manual_seed(18)
    leaves = torch.randint(0, 1 << 62, (n, leaf_len), dtype=torch.int64, device="cuda", generator=g)
    total = 4 * (2 * n - (1 << cap_height))
    layers = torch.zeros(total, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.merkle_tree_build(leaves.data_ptr(), leaf_len, cap_height, n_leaves=n, layers_out=layers.data_ptr(), mem=svb.MEM_DEVICE)
    ctx.synchronize()
    depth = log_n - cap_height
    m = 1 << 16
    idx = torch.randint(0, n, (m,), dtype=torch.int64, device="cuda", generator=g)
    lw = (leaf_len + 3) & ~3
    paths = torch.zeros((m, lw + 4 * depth), dtype=torch.int64, device="cuda")
    paths[:, :leaf_len] = leaves[idx]
    off, width, node = 0, n, idx.clone()
    for lvl in range(depth):
        layer = layers[off:off + 4 * width].view(width, 4)
        paths[:, lw + 4 * lvl: lw + 4 * lvl + 4] = layer[node ^ 1]
        off += 4 * width; width >>= 1; node >>= 1
    cap = layers[off:off + 4 * width].contiguous()
    assert width == 1 << cap_height
    bad = torch.arange(0, m, 4, device="cuda")
    col = lw + 4 * (torch.arange(bad.numel(), device="cuda") % depth) + 1
    paths[bad, col] ^= 1
    ok = torch.zeros(m, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.merkle_verify_batch(leaf_len, depth, cap_height, paths.data_ptr(), idx.data_ptr(), cap.data_ptr(), ok.data_ptr(), n=m,
                            mem=svb.MEM_DEVICE)
    ctx.synchronize()
    want = torch.ones(m, dtype=torch.uint8, device="cuda")
    want[bad] = 0
    assert torch.equal(ok, want)


@pytest.mark.parametrize("polys,num_zs,hiding,num_challenges", [((6, 11, 4, 3), 2, False, 2), ((6, 11, 4, 3), 1, True, 1),
                                                                 ((1, 2, 3, 4), 3, False, 3), ((9, 1, 8, 1), 2, True, 2)])
def test_fri_unusual_oracle_widths(svb, orc, ctx, polys, num_zs, hiding, num_challenges):
    """Oracle widths other than the recursion-circuit defaults: leaves of <= 4 words inside the FRI path
    (hash_or_noop: the leaf IS the digest), one-column oracles, a different number of Z polynomials /
    challenges (the transcript squeezes num_challenges betas, gammas and alphas before zeta)."""
    params = tiny_params(svb, hiding=hiding, cap=1, degree_bits=6, oracle_num_polys=polys, num_zs=num_zs)
    L = svb.api.make_layout(params)
    n = 33
    recs = svb.synth_proofs(params, n, seed=sum(polys), n_circuits=1, num_challenges=num_challenges)
    bad = corrupt(recs, L, np.random.default_rng(12), every=4)
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    oshape = orc.shape_from(params.to_shape())
    want = orc.fri_verify_batch(oshape, recs, nthreads=4)
    assert (bm == want).all()
    for i in range(n):
        ok, code, q = orc.fri_verify(oshape, recs[i])
        assert int(ff[i]) == (0 if ok else ((max(q, 0) << 8) | code))
        assert bit(bm, i) == (0 if i in bad else 1), (i, bad.get(i), code)
    # the device transcript with this num_challenges reproduces the prover's challenges
    cd, ph = svb.synth_public_inputs(params, 8, seed=sum(polys), n_circuits=1)
    fresh = svb.synth_proofs(params, 8, seed=sum(polys), n_circuits=1, num_challenges=num_challenges)
    dev = _clear_challenges(fresh, L, params)
    ctx.fri_challenges_batch(params, dev, cd[0], ph, num_challenges=num_challenges)
    assert (dev == fresh).all()


def test_device_transcript_thread_per_proof_kernel(svb):
    """The thread-per-proof transcript kernel (used above 8 192 proofs per call) on a small batch, forced with
    SVB_FS_COOP=0 in a child process (the choice is read once per process)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import stark_verifier_b200 as svb
from common import tiny_params
params = tiny_params(svb, cap=2, degree_bits=7)
recs = svb.synth_proofs(params, 40, seed=77, n_circuits=1)
cd, ph = svb.synth_public_inputs(params, 40, seed=77, n_circuits=1)
L = svb.api.make_layout(params)
dev = recs.copy()
for off, n in ((L.off_alpha, 2), (L.off_betas, 4), (L.off_pow_response, 1), (L.off_indices, 6), (L.off_zeta, 2), (L.off_zeta_next, 2)):
    dev[:, off:off + n] = 0
ctx = svb.Context(0)
ctx.fri_challenges_batch(params, dev, cd[0], ph)
assert (dev == recs).all()
print("ok")
''' % (ROOT_DIR, os.path.join(ROOT_DIR, "tests"))
    env = dict(os.environ, SVB_FS_COOP="0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-1500:]
