"""Wire format, CPU side: the product's table-driven packer / unpacker (stark-verifier_b200/csrc/wire.hpp,
wire_host.cpp -- the unpacker runs the very word function the device gather kernel runs) against the oracle's
independent cursor-style reader / writer (oracle/wire.c), a pure-Python sequential writer, the committed golden blob,
and the round trip bytes -> record -> transcript -> verdict."""
import os
import struct

import numpy as np
import pytest

from common import P, bit, tiny_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = [dict(), dict(hiding=True), dict(cap=0, degree_bits=8, rate_bits=2), dict(cap=4, degree_bits=6),
          dict(degree_bits=5, queries=3),                     # no reduction steps
          dict(degree_bits=6, rate_bits=1, cap=2, queries=2)]  # last step tree has depth 0 (length byte 0)


def make(svb, kw, n=5, n_pi=5, seed=11):
    params = tiny_params(svb, **kw)
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    L = svb.api.make_layout(params)
    recs = svb.synth_proofs(params, n, seed=seed, n_circuits=2, nthreads=4)
    rng = np.random.default_rng(seed)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    return params, common, L, recs, pis


def python_writer(params, common, L, rec, pis):
    """write_proof_with_public_inputs, statement by statement (third restatement, no shared code)."""
    out = bytearray()
    f = lambda words: out.extend(struct.pack("<%dQ" % len(words), *[int(w) for w in words]))
    capw = 4 * L.ncap
    for k in (1, 2, 3):                                   # wires_cap, zs_partial_products_cap, quotient_polys_cap
        f(rec[L.off_init_caps + k * capw: L.off_init_caps + (k + 1) * capw])
    c = common
    o0 = L.off_open0
    sizes = [c.num_constants, c.num_routed_wires, c.num_wires, c.num_challenges]
    for s in sizes:                                       # constants, plonk_sigmas, wires, plonk_zs
        f(rec[o0: o0 + 2 * s]); o0 += 2 * s
    f(rec[L.off_open1: L.off_open1 + 2 * c.num_challenges])   # plonk_zs_next
    for s in (c.num_challenges * c.num_partial_products, c.num_challenges * c.quotient_degree_factor):
        f(rec[o0: o0 + 2 * s]); o0 += 2 * s
    steps = len(params.reduction_arity_bits)
    f(rec[L.off_step_caps: L.off_step_caps + steps * capw])
    for q in range(params.config.num_query_rounds):
        qb = L.header_words + q * L.query_words
        for k in range(4):
            f(rec[qb + L.q_off_init_evals[k]: qb + L.q_off_init_evals[k] + L.leaf_len[k]])
            out.append(L.init_depth)
            f(rec[qb + L.q_off_init_sibs[k]: qb + L.q_off_init_sibs[k] + 4 * L.init_depth])
        for i in range(steps):
            f(rec[qb + L.q_off_step_evals[i]: qb + L.q_off_step_evals[i] + 4])
            out.append(L.step_depth[i])
            f(rec[qb + L.q_off_step_sibs[i]: qb + L.q_off_step_sibs[i] + 4 * L.step_depth[i]])
    f(rec[L.off_final_poly: L.off_final_poly + 2 * params.final_poly_len()])
    f([rec[L.off_pow_witness]])
    f(pis)
    return np.frombuffer(bytes(out), dtype=np.uint8)


def expected_records(recs, L, cap):
    exp = recs.copy()
    exp[:, L.off_alpha:L.header_words] = 0        # challenges are not part of a proof
    exp[:, L.off_init_caps:L.off_init_caps + 4 * L.ncap] = cap
    return exp


@pytest.mark.parametrize("kw", SHAPES)
def test_pack_matches_oracle_and_python_writers(svb, orc, kw):
    params, common, L, recs, pis = make(svb, kw)
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    blob = svb.wire_pack(common, recs, pis)
    assert blob.shape[1] == svb.wire_proof_bytes(common) == orc.wire_proof_bytes(oshape, ocommon)
    for i in range(recs.shape[0]):
        assert (blob[i] == orc.wire_write_proof(oshape, ocommon, recs[i], pis[i])).all()
        assert (blob[i] == python_writer(params, common, L, recs[i], pis[i])).all()


def test_shape_sizes():
    """Bytes of one proof: algorithmic payload (minus the verifier key's cap and the challenges that are
    not in a proof) + one length byte per Merkle proof + public inputs."""
    import stark_verifier_b200 as svb
    A = svb.api.make_layout(svb.SHAPE_A)
    cd = svb.CommonData.for_params(svb.SHAPE_A, num_public_inputs=4)
    want = 28 * A.algo_bytes_per_query + A.algo_bytes_shared - 32 * A.ncap + 28 * (4 + 7) + 8 * 4
    assert svb.wire_proof_bytes(cd) == want == 156812
    s = svb.shape_from_common(cd)
    assert bytes(s) == bytes(svb.SHAPE_A.to_shape())


@pytest.mark.parametrize("kw", SHAPES)
@pytest.mark.parametrize("misalign", [0, 3])
def test_unpack_matches_oracle_reader(svb, orc, kw, misalign):
    params, common, L, recs, pis = make(svb, kw, n=6)
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    blob = svb.wire_pack(common, recs, pis)
    vk_cap = np.arange(1, 4 * L.ncap + 1, dtype=np.uint64)      # the verifier key's cap replaces init_caps[0]
    # every alignment of the proofs inside the aligned words the gather reads
    buf = np.zeros(blob.size + 8, dtype=np.uint8)
    buf[misalign:misalign + blob.size] = blob.reshape(-1)
    r2, pih, pi2, mal = svb.wire_unpack_batch(common, vk_cap, buf[misalign:misalign + blob.size], nthreads=3)
    assert (r2 == expected_records(recs, L, vk_cap)).all()
    assert (pi2 == pis).all() and not mal.any()
    for i in range(recs.shape[0]):
        rc, orec, opis, opih = orc.wire_read_proof(oshape, ocommon, vk_cap, blob[i])
        assert rc == 0
        assert (orec == r2[i]).all() and (opis == pis[i]).all()
        assert (opih == pih[i]).all() and (opih == orc.hash_no_pad(pis[i])).all()
        assert (svb.public_inputs_hash(pis[i]) == pih[i]).all()


def test_unpack_with_padded_stride(svb):
    params, common, L, recs, pis = make(svb, dict(), n=4)
    blob = svb.wire_pack(common, recs, pis)
    nb = blob.shape[1]
    wide = np.full((4, nb + 13), 0xAB, dtype=np.uint8)
    wide[:, :nb] = blob
    cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    r2, _, pi2, mal = svb.wire_unpack_batch(common, cap, wide.reshape(-1), n_proofs=4, stride=nb + 13)
    assert (r2 == expected_records(recs, L, cap)).all() and (pi2 == pis).all() and not mal.any()


def test_malformed_flags(svb, orc):
    params, common, L, recs, pis = make(svb, dict(), n=8)
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    blob = svb.wire_pack(common, recs, pis)
    cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    # positions of the Merkle-proof length bytes: wherever the byte differs when the writer is told another depth
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    first_len_byte = q0 + 8 * L.leaf_len[0]
    assert blob[0, first_len_byte] == L.init_depth
    blob[2, first_len_byte] += 1                       # first path of the first query round
    qbytes = (blob.shape[1] - q0 - 16 * params.final_poly_len() - 8 - 8 * common.num_public_inputs) // params.config.num_query_rounds
    last = q0 + (params.config.num_query_rounds - 1) * qbytes + qbytes - 32 * L.step_depth[len(params.reduction_arity_bits) - 1] - 1
    assert blob[5, last] == L.step_depth[len(params.reduction_arity_bits) - 1]
    blob[5, last] = 0xFF                               # last step path of the last query round
    blob[6, -8:] = np.frombuffer(struct.pack("<Q", P), dtype=np.uint8)   # a public input >= p
    _, _, _, mal = svb.wire_unpack_batch(common, cap, blob.reshape(-1), nthreads=2)
    assert list(mal) == [0, 0, 1, 0, 0, 1, 1, 0]
    for i in range(8):
        rc = orc.wire_read_proof(oshape, ocommon, cap, blob[i])[0]
        assert rc == int(mal[i])
    # wrong length cannot be framed at all
    assert orc.wire_read_proof(oshape, ocommon, cap, blob[0][:-1])[0] < 0


def test_shape_common_mismatch_is_an_error(svb):
    params = tiny_params(svb)
    common = svb.CommonData.for_params(params)
    common.num_wires += 1
    with pytest.raises(svb.SvError):
        svb.wire_proof_bytes(common)


@pytest.mark.parametrize("kw", [dict(), dict(hiding=True, cap=0, degree_bits=8, rate_bits=2)])
def test_bytes_to_verdict_round_trip(svb, orc, kw):
    """bytes -> record + public-inputs hash -> transcript (host) -> oracle verdict: a proof that went through the wire
    format verifies exactly like the record it came from, and a flipped wire byte rejects it."""
    params = tiny_params(svb, **kw)
    L = svb.api.make_layout(params)
    n, n_pi = 6, 3
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    rng = np.random.default_rng(5)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    # proofs whose transcripts are bound to hash(public inputs): build with the synthetic prover's own pi hashes,
    # then re-derive the challenges for OUR public inputs on both sides of the wire
    recs = svb.synth_proofs(params, n, seed=21, n_circuits=1, nthreads=4)
    cds, _ = svb.synth_public_inputs(params, n, seed=21, n_circuits=1)
    cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    blob = svb.wire_pack(common, recs, pis)
    blob[4, 3 * 32 * L.ncap + 5] ^= 1                   # corrupt one opening of proof 4
    r2, pih, _, mal = svb.wire_unpack_batch(common, cap, blob.reshape(-1))
    assert not mal.any()
    oshape = orc.shape_from(params.to_shape())
    for i in range(n):
        svb.fri_challenges(params, r2[i], cds[0], pih[i])
    # the prover bound its proofs to other public-input hashes, so the re-derived challenges differ from the ones the
    # proof was built for: what must hold is that the wire path and the record path agree bit for bit
    direct = recs.copy()
    direct[4, L.off_open0 + 0] ^= np.uint64(1 << 40)    # same corruption: byte 5 of the first opening word
    for i in range(n):
        svb.fri_challenges(params, direct[i], cds[0], svb.public_inputs_hash(pis[i]))
    assert (direct == r2).all()
    assert (orc.fri_verify_batch(oshape, r2) == orc.fri_verify_batch(oshape, direct)).all()


def test_bytes_to_verdict_accepts_valid_proofs(svb, orc):
    """Proofs bound to hash(public inputs): serialised, unpacked, challenges re-derived from the unpacked public
    inputs -- the records come back bit for bit (challenges included) and the oracle accepts them; a proof whose
    public inputs were tampered with on the wire derives other challenges and is rejected."""
    params = tiny_params(svb, hiding=True)
    L = svb.api.make_layout(params)
    n, n_pi = 8, 9
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    rng = np.random.default_rng(8)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    pih = np.stack([svb.public_inputs_hash(pis[i]) for i in range(n)])
    recs = svb.synth_proofs(params, n, seed=33, n_circuits=1, nthreads=4, pi_hashes=pih)
    cds, _ = svb.synth_public_inputs(params, n, seed=33, n_circuits=1)
    oshape = orc.shape_from(params.to_shape())
    assert all(orc.fri_verify(oshape, r)[0] for r in recs)
    blob = svb.wire_pack(common, recs, pis)
    blob[6, -3] ^= 1                                     # one bit of the last public input of proof 6
    cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    r2, pih2, _, mal = svb.wire_unpack_batch(common, cap, blob.reshape(-1), nthreads=2)
    assert not mal.any()
    for i in range(n):
        svb.fri_challenges(params, r2[i], cds[0], pih2[i])
    good = [i for i in range(n) if i != 6]
    assert (r2[good] == recs[good]).all() and (pih2[good] == pih[good]).all()
    bm = orc.fri_verify_batch(oshape, r2)
    assert [bit(bm, i) for i in range(n)] == [1, 1, 1, 1, 1, 1, 0, 1]


def test_golden_wire_blob(svb, orc):
    """tests/golden/wire_small.npz (tools/gen_golden.py): committed bytes, records and public-input hashes."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "wire_small.npz"))
    params = tiny_params(svb, **{k: int(v) for k, v in zip(g["param_names"], g["param_values"])})
    common = svb.CommonData.for_params(params, num_public_inputs=int(g["num_public_inputs"]))
    r2, pih, pi2, mal = svb.wire_unpack_batch(common, g["vk_cap"], g["blob"].reshape(-1))
    assert (r2 == g["records"]).all() and (pih == g["pi_hashes"]).all() and (pi2 == g["public_inputs"]).all()
    assert list(mal) == list(g["malformed"])
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    for i in range(r2.shape[0]):
        rc, orec, _, opih = orc.wire_read_proof(oshape, ocommon, g["vk_cap"], g["blob"][i])
        assert rc == int(g["malformed"][i]) and (orec == g["records"][i]).all() and (opih == g["pi_hashes"][i]).all()


def test_random_bytes_never_disagree_with_the_oracle_reader(svb, orc):
    """Fuzz: arbitrary bytes of the right length (random blobs, and valid proofs with random byte flips) unpack to the same
    record, public inputs, hash and malformed flag through the product's tables and through the oracle's cursor reader."""
    params, common, L, recs, pis = make(svb, dict(hiding=True, cap=1, degree_bits=6, rate_bits=2, queries=3), n=4, n_pi=10)
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    rng = np.random.default_rng(2024)
    blob = svb.wire_pack(common, recs, pis)
    nb = blob.shape[1]
    cases = [rng.integers(0, 256, size=nb, dtype=np.uint8) for _ in range(6)]
    cases.append(np.full(nb, 0xFF, dtype=np.uint8))
    for i in range(24):
        b = blob[i % 4].copy()
        for pos in rng.integers(0, nb, size=int(rng.integers(1, 40))):
            b[pos] = rng.integers(0, 256)
        cases.append(b)
    all_bytes = np.stack(cases)
    cap = rng.integers(0, P, size=4 * L.ncap, dtype=np.uint64)
    r2, pih, pi2, mal = svb.wire_unpack_batch(common, cap, all_bytes.reshape(-1), nthreads=3)
    for i, b in enumerate(cases):
        rc, orec, opis, opih = orc.wire_read_proof(oshape, ocommon, cap, b)
        assert rc == int(mal[i]) and (orec == r2[i]).all() and (opis == pi2[i]).all() and (opih == pih[i]).all(), i
    assert mal[:7].all()          # random bytes never carry the right length bytes
    # unpack -> pack is the identity on the bytes whenever the length bytes are right
    for i in np.flatnonzero(mal == 0):
        again = svb.wire_pack(common, r2[i:i + 1], pi2[i:i + 1])[0]
        assert (again == cases[i]).all()


def test_host_entry_points_report_bad_arguments(svb):
    """Error convention of the new host entry points: < 0 for bad arguments, never a crash, never a silent success."""
    import ctypes
    L = svb.lib()
    params = tiny_params(svb)
    common = svb.CommonData.for_params(params, num_public_inputs=2)
    s, c = params.to_shape(), common.to_c()
    nb = svb.wire_proof_bytes(common)
    cap = np.zeros(4 << s.cap_height, dtype=np.uint64)
    blob = np.zeros(3 * nb, dtype=np.uint8)
    recs = np.zeros((3, svb.api.make_layout(params).record_words), dtype=np.uint64)
    p = lambda a: a.ctypes.data
    assert L.sv_wire_unpack_batch(ctypes.byref(s), ctypes.byref(c), p(cap), p(blob), nb - 1, 3, p(recs), None, None, None, 1) == -4   # stride < proof
    assert L.sv_wire_unpack_batch(ctypes.byref(s), ctypes.byref(c), None, p(blob), nb, 3, p(recs), None, None, None, 1) == -1          # no verifier key
    assert L.sv_wire_unpack_batch(None, ctypes.byref(c), p(cap), p(blob), nb, 3, p(recs), None, None, None, 1) == -1
    assert L.sv_wire_unpack_batch(ctypes.byref(s), ctypes.byref(c), p(cap), p(blob), nb, 0, p(recs), None, None, None, 1) == 0            # empty batch
    assert L.sv_wire_pack(ctypes.byref(s), ctypes.byref(c), p(recs), None, p(blob)) == -1                                               # public inputs missing
    assert L.sv_wire_proof_bytes(None, ctypes.byref(c)) == 0
    assert L.sv_public_inputs_hash(None, 3, p(cap)) == -1
    out = svb.FriShape()
    assert L.sv_fri_shape_from_common(ctypes.byref(c), 4, 3, 1, 5, 2, 9, None, 0, 0, ctypes.byref(out)) == -2                                 # more steps than degree bits
    assert L.sv_ntt_host(0, 1, p(recs), 0, 1) == -2 and L.sv_ntt_host(3, 1, None, 0, 1) == -1
    assert L.sv_plonk_gate_from_id(None, None) == -1
    assert L.sv_plonk_check_host(ctypes.byref(s), None, 1, p(recs), p(cap), p(cap), p(cap), 1) == -1


def test_random_shapes_round_trip_against_the_oracle(svb, orc):
    """40 seeded random circuit shapes (widths, depths, cap heights, salting, public inputs, both step-count regimes) with
    random record contents: product pack == oracle write, product unpack == oracle read, pack(unpack(bytes)) == bytes."""
    rng = np.random.default_rng(777)
    for case in range(40):
        degree_bits = int(rng.integers(3, 10))
        rate_bits = int(rng.integers(1, 4))
        cap = int(rng.integers(0, min(5, degree_bits + rate_bits - max(0, degree_bits - 5)) + 1))
        nch = int(rng.integers(1, 4))
        nc, nr = int(rng.integers(1, 8)), int(rng.integers(1, 30))
        nw = nr + int(rng.integers(0, 10))
        npp, qdf = int(rng.integers(0, 5)), int(rng.integers(1, 9))
        widths = (nc + nr, nw, nch * (1 + npp), nch * qdf)
        params = svb.api._params(degree_bits, rate_bits, cap, int(rng.integers(0, 8)), int(rng.integers(1, 9)),
                                 hiding=bool(rng.integers(0, 2)), oracle_num_polys=widths, num_zs=nch)
        common = svb.CommonData(params, nc, nr, nw, nch, npp, qdf, int(rng.integers(0, 12)))
        L = svb.api.make_layout(params)
        oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
        nb = svb.wire_proof_bytes(common)
        assert nb == orc.wire_proof_bytes(oshape, ocommon)
        n = 3
        pis = rng.integers(0, P, size=(n, common.num_public_inputs), dtype=np.uint64)
        recs = rng.integers(0, P, size=(n, L.record_words), dtype=np.uint64)
        blob = svb.wire_pack(common, recs, pis)
        cap_words = rng.integers(0, P, size=4 * L.ncap, dtype=np.uint64)
        r2, pih, pi2, mal = svb.wire_unpack_batch(common, cap_words, blob.reshape(-1), nthreads=2)
        assert not mal.any() and (pi2 == pis).all(), case
        assert (svb.wire_pack(common, r2, pi2) == blob).all(), case
        for i in range(n):
            assert (blob[i] == orc.wire_write_proof(oshape, ocommon, recs[i], pis[i])).all(), case
            rc, orec, opis, opih = orc.wire_read_proof(oshape, ocommon, cap_words, blob[i])
            assert rc == 0 and (orec == r2[i]).all() and (opih == pih[i]).all(), case
        # what unpack gives back: the proof's own fields, the verifier key's cap, zeros everywhere else
        again = r2.copy()
        again[:, L.off_init_caps:L.off_init_caps + 4 * L.ncap] = recs[:, L.off_init_caps:L.off_init_caps + 4 * L.ncap]
        diff = np.flatnonzero((again != recs).any(axis=0))
        assert all(r2[0, w] == 0 for w in diff), case
