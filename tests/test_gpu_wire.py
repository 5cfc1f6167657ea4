"""GPU side of the wire format: the gather kernel and the public-input hash kernel against the CPU
unpacker (which tests/test_wire_format.py checks against the oracle's independent reader), and the bytes -> verdict
pipeline sv_verify_proofs_wire against the oracle's verdict on the same proofs."""
import os

import numpy as np
import pytest

from common import P, bit, tiny_params

pytestmark = pytest.mark.gpu


def bound_proofs(svb, params, n, n_pi, seed, n_circuits=1):
    """n valid proofs whose transcripts are bound to hash(public inputs), with those public inputs."""
    rng = np.random.default_rng(seed)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    pih = np.stack([svb.public_inputs_hash(pis[i]) for i in range(n)]) if n else np.zeros((0, 4), dtype=np.uint64)
    recs = svb.synth_proofs(params, n, seed=seed, n_circuits=n_circuits, pi_hashes=pih)
    cds, _ = svb.synth_public_inputs(params, n, seed=seed, n_circuits=n_circuits)
    return recs, pis, pih, cds


@pytest.mark.parametrize("kw,n_pi", [(dict(), 5), (dict(hiding=True, cap=0, degree_bits=8, rate_bits=2), 17),
                                     (dict(degree_bits=5, queries=3), 0), (dict(degree_bits=6, rate_bits=1, cap=2, queries=2), 8)])
@pytest.mark.parametrize("pad", [0, 5])
def test_unpack_kernel_matches_cpu_unpacker(svb, ctx, kw, n_pi, pad):
    params = tiny_params(svb, **kw)
    L = svb.api.make_layout(params)
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    n = 37
    recs, pis, pih, _ = bound_proofs(svb, params, n, n_pi, seed=3)
    blob = svb.wire_pack(common, recs, pis)
    nb = blob.shape[1]
    # Merkle-proof length bytes: first path of the first query round of proof 4, a public input >= p in proof 9
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    blob[4, q0 + 8 * L.leaf_len[0]] ^= 0x10
    if n_pi:
        blob[9, -8:] = np.frombuffer(np.uint64(P + 2).tobytes(), dtype=np.uint8)
    wide = np.full((n, nb + pad), 0x5A, dtype=np.uint8)
    wide[:, :nb] = blob
    vk_cap = np.arange(7, 7 + 4 * L.ncap, dtype=np.uint64)
    want_r, want_h, _, want_m = svb.wire_unpack_batch(common, vk_cap, wide.reshape(-1), n_proofs=n, stride=nb + pad, nthreads=4)
    got_r, got_h, got_m = ctx.wire_unpack_batch(common, vk_cap, wide.reshape(-1), n_proofs=n, stride=nb + pad)
    assert (got_r == want_r).all()
    assert (got_h == want_h).all()
    assert (got_m.astype(np.uint8) == want_m).all()
    assert want_m[4] == 1 and (want_m[9] == 1) == (n_pi > 0) and want_m.sum() == (2 if n_pi else 1)


def test_unpack_kernel_device_memory(svb, ctx):
    import torch
    params = tiny_params(svb, cap=3, degree_bits=8)
    L = svb.api.make_layout(params)
    common = svb.CommonData.for_params(params, num_public_inputs=6)
    n = 50
    recs, pis, pih, _ = bound_proofs(svb, params, n, 6, seed=4)
    blob = svb.wire_pack(common, recs, pis)
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    want_r, want_h, _, want_m = svb.wire_unpack_batch(common, vk_cap, blob.reshape(-1), nthreads=4)
    flat = np.zeros((blob.size + 15) // 8 * 8, dtype=np.uint8)       # readable up to the next multiple of 8
    flat[:blob.size] = blob.reshape(-1)
    d_blob = torch.from_numpy(flat).cuda()
    d_rec = torch.zeros(n * L.record_words, dtype=torch.int64, device="cuda")
    d_pih = torch.zeros(n * 4, dtype=torch.int64, device="cuda")
    d_mal = torch.ones(n, dtype=torch.int32, device="cuda")          # the call zeroes the flags itself
    torch.cuda.synchronize()
    ctx.wire_unpack_batch(common, vk_cap, d_blob.data_ptr(), n_proofs=n, records_out=d_rec.data_ptr(),
                          pi_hashes_out=d_pih.data_ptr(), malformed_out=d_mal.data_ptr(), mem=svb.MEM_DEVICE)
    ctx.synchronize()
    assert (d_rec.cpu().numpy().view(np.uint64).reshape(n, -1) == want_r).all()
    assert (d_pih.cpu().numpy().view(np.uint64).reshape(n, 4) == want_h).all()
    assert not d_mal.cpu().numpy().any() and not want_m.any()


@pytest.mark.parametrize("kw,n_pi,kind", [(dict(cap=3, degree_bits=8), 7, 0), (dict(hiding=True, cap=0, degree_bits=7, rate_bits=2), 3, 0),
                                          (dict(cap=1, degree_bits=6), 4, 1)])
def test_verify_proofs_wire_matches_oracle(svb, orc, ctx, kw, n_pi, kind):
    params = tiny_params(svb, hash_kind=kind, **kw)
    L = svb.api.make_layout(params)
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    n = 70 if kind == 0 else 20
    recs, pis, pih, cds = bound_proofs(svb, params, n, n_pi, seed=6)
    oshape = orc.shape_from(params.to_shape())
    assert all(bit(orc.fri_verify_batch(oshape, recs, nthreads=4), i) for i in range(n))
    blob = svb.wire_pack(common, recs, pis)
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    blob[2, q0 + 8 * L.leaf_len[0] + 1 + 5] ^= 1          # a sibling byte (query round 0, oracle 0)
    blob[5, 3 * 32 * L.ncap + 9] ^= 2                     # an opening: changes the transcript too
    blob[8, -2] ^= 1                                      # a public input: other challenges
    blob[11, q0 + 8 * L.leaf_len[0]] += 1                 # malformed: wrong Merkle-proof length byte
    blob[13, q0 + 3] ^= 4                                 # a leaf evaluation
    # expectation from the CPU side: unpack, host transcript, oracle verdict
    r2, pih2, _, mal = svb.wire_unpack_batch(common, vk_cap, blob.reshape(-1), nthreads=4)
    for i in range(n):
        svb.fri_challenges(params, r2[i], cds[0], pih2[i])
    want = orc.fri_verify_batch(oshape, r2, nthreads=4)
    expect = [bit(want, i) and not mal[i] for i in range(n)]
    assert [i for i in range(n) if not expect[i]] == [2, 5, 8, 11, 13]
    bm, ff = ctx.verify_proofs_wire(common, vk_cap, cds[0], blob.reshape(-1), want_fail=True)
    assert [bit(bm, i) for i in range(n)] == [int(e) for e in expect]
    assert ff[11] == svb.FAIL_MALFORMED and all(ff[i] == 0 for i in range(n) if expect[i])
    # the same verdicts as the record path with the device transcript
    bm2 = ctx.fri_verify_batch_fs(params, r2, cds[0], pih2)
    for i in range(n):
        if i != 11:
            assert bit(bm2, i) == bit(bm, i)
    # idempotent
    assert (ctx.verify_proofs_wire(common, vk_cap, cds[0], blob.reshape(-1)) == bm).all()


@pytest.mark.parametrize("chunk_mb,n", [(1, 16 * 15 + 5), (6, 560), (4, 645)])
def test_verify_proofs_wire_many_chunks(svb, orc, ctx, chunk_mb, n):
    """More chunks than staging buffers (the ring wraps), a ragged last chunk, pinned host memory; (6, 560): chunks of 224, 192,
    96, 48 proofs -- the ramp-down of the chunk schedule, with one chunk straddling the two transcript parts (boundary at 448);
    (4, 645): 128, 128, 128, 128, 96, 37."""
    import ctypes
    params = tiny_params(svb)
    L = svb.api.make_layout(params)
    common = svb.CommonData.for_params(params, num_public_inputs=2)
    base_n = 16
    recs, pis, pih, cds = bound_proofs(svb, params, base_n, 2, seed=9)
    blob16 = svb.wire_pack(common, recs, pis)
    nb = blob16.shape[1]
    p = ctypes.c_void_p()
    assert svb.lib().sv_host_alloc(n * nb, ctypes.byref(p)) == 0
    try:
        host = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n, nb))
        for i in range(n):
            host[i] = blob16[i % base_n]
        q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
        bad = list(range(3, n, 29))
        for i in bad:
            host[i, q0 + 8 * L.leaf_len[0] + 1 + 32 + 4] ^= 1            # a sibling of oracle 0 in query round 0
        vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
        r2, pih2, _, mal = svb.wire_unpack_batch(common, vk_cap, host.reshape(-1), nthreads=4)
        for i in range(n):
            svb.fri_challenges(params, r2[i], cds[0], pih2[i])
        want = orc.fri_verify_batch(orc.shape_from(params.to_shape()), r2, nthreads=4)
        assert not mal.any() and [i for i in range(n) if not bit(want, i)] == bad
        old = {k: os.environ.get(k) for k in ("SVB_CHUNK_MB", "SVB_RAMP")}
        os.environ["SVB_CHUNK_MB"] = str(chunk_mb)        # 1 MiB: 32 proofs per chunk -> 8 chunks over 6 buffers
        os.environ["SVB_RAMP"] = "1"                      # the ramp-down is off by default on the wire path
        try:
            bm = ctx.verify_proofs_wire(common, vk_cap, cds[0], p.value, n_proofs=n)
            os.environ["SVB_RAMP"] = "0"
            bm0 = ctx.verify_proofs_wire(common, vk_cap, cds[0], p.value, n_proofs=n)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        assert (bm == bm0).all()
        assert [i for i in range(n) if not bit(bm, i)] == bad
        assert int(bm[-1]) >> (n & 31) == 0               # bits past the batch stay clear
    finally:
        svb.lib().sv_host_free(p)


def test_verify_proofs_wire_empty_and_errors(svb, ctx):
    params = tiny_params(svb)
    common = svb.CommonData.for_params(params, num_public_inputs=1)
    L = svb.api.make_layout(params)
    cap = np.zeros(4 * L.ncap, dtype=np.uint64)
    bm = ctx.verify_proofs_wire(common, cap, np.zeros(4, dtype=np.uint64), np.zeros(0, dtype=np.uint8), n_proofs=0)
    assert bm.size == 0
    bad = svb.CommonData.for_params(params, num_public_inputs=1)
    bad.num_wires += 1
    with pytest.raises(svb.SvError):
        ctx.verify_proofs_wire(bad, cap, np.zeros(4, dtype=np.uint64), np.zeros(64, dtype=np.uint8), n_proofs=1, stride=64)
