// Host side of the wire format: table construction, packer, CPU unpacker, public-inputs hash.
// See wire.hpp for the format and for the word-level function shared with the device gather kernel.
#include "wire.hpp"
#include "host_util.hpp"

#include <cstdlib>

namespace svb {

static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, "wire words are little-endian u64; so is every host CUDA runs on");

int make_wire_map(const sv_fri_shape& s, const sv_plonk_common& c, WireMap& M) {
    sv_fri_layout L;
    if (make_layout(s, L)) return -1;
    // CommonData::fri_oracles / fri_zs_polys (types/common_data.rs:153-221)
    if (s.oracle_num_polys[0] != c.num_constants + c.num_routed_wires) return -2;
    if (s.oracle_num_polys[1] != c.num_wires) return -2;
    if (s.oracle_num_polys[2] != c.num_challenges * (1 + c.num_partial_products)) return -2;
    if (s.oracle_num_polys[3] != c.num_challenges * c.quotient_degree_factor) return -2;
    if (s.num_zs != c.num_challenges || c.num_challenges == 0) return -2;
    WireDims& d = M.d;
    d.header_words = L.header_words;
    d.query_words = L.query_words;
    d.record_words = L.record_words;
    d.num_queries = s.num_query_rounds;
    d.num_public_inputs = c.num_public_inputs;
    d.n_chk = 4 + s.num_steps;
    M.hdr_src.assign(L.header_words, WIRE_ZERO);
    M.q_src.assign(L.query_words, WIRE_ZERO);
    M.chk.assign(d.n_chk, 0);

    uint64_t o = 0;  // byte cursor, in the order of write_proof_with_public_inputs
    auto run = [&](std::vector<u32>& tab, u32 dst_word, uint64_t words) {
        for (uint64_t j = 0; j < words; j++) tab[dst_word + j] = (u32)(o + 8 * j);
        o += 8 * words;
    };
    const u32 cap_words = L.ncap * 4;
    // init_caps[0] comes from the verifier key (VerificationKeyValues.constants_sigmas_cap), not from the proof
    for (u32 j = 0; j < cap_words; j++) M.hdr_src[L.off_init_caps + j] = WIRE_VK_FLAG | j;
    // wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap (ProofValues, types/proof.rs:380-387)
    run(M.hdr_src, L.off_init_caps + cap_words, 3ull * cap_words);
    // OpeningSetValues order (types/proof.rs:34-43): constants, plonk_sigmas, wires, plonk_zs | plonk_zs_next |
    // partial_products, quotient_polys; FRI batch 0 is the same list without plonk_zs_next (types/assigned.rs:26-37)
    const u32 head = c.num_constants + c.num_routed_wires + c.num_wires + c.num_challenges;
    run(M.hdr_src, L.off_open0, 2ull * head);
    run(M.hdr_src, L.off_open1, 2ull * c.num_challenges);
    run(M.hdr_src, L.off_open0 + 2 * head, 2ull * (L.n0 - head));
    run(M.hdr_src, L.off_step_caps, (uint64_t)s.num_steps * cap_words);
    d.query_base = (u32)o;
    {
        const uint64_t o_keep = o;
        o = 0;  // offsets inside one query round (FriQueryRoundValues, types/proof.rs:217-222)
        u32 n = 0;
        for (int k = 0; k < 4; k++) {
            run(M.q_src, L.q_off_init_evals[k], L.leaf_len[k]);
            M.chk[n++] = (u32)(o << 8) | L.init_depth;
            o += 1;
            run(M.q_src, L.q_off_init_sibs[k], 4ull * L.init_depth);
        }
        for (u32 i = 0; i < s.num_steps; i++) {
            run(M.q_src, L.q_off_step_evals[i], 2ull << L.step_arity_bits[i]);
            M.chk[n++] = (u32)(o << 8) | L.step_depth[i];
            o += 1;
            run(M.q_src, L.q_off_step_sibs[i], 4ull * L.step_depth[i]);
        }
        if (o >= (1u << 24)) return -3;
        d.query_bytes = (u32)o;
        o = o_keep + (uint64_t)s.num_query_rounds * d.query_bytes;
    }
    if (o >= (1ull << 31)) return -3;
    run(M.hdr_src, L.off_final_poly, 2ull * s.final_poly_len);
    run(M.hdr_src, L.off_pow_witness, 1);
    d.pi_off = (u32)o;
    o += 8ull * c.num_public_inputs;
    if (o >= (1ull << 31)) return -3;
    d.proof_bytes = (u32)o;
    return 0;
}

// One proof through the shared word function.  base: 8-byte aligned, off: where the proof starts.
static void unpack_one(const WireMap& M, const u64* vk_cap, const u64* base, size_t off, u64* rec, u64* pi_hash, u64* pis,
                       unsigned char* malformed) {
    bool bad = false;
    for (u32 w = 0; w < M.d.record_words; w++)
        rec[w] = wire_record_word(M.d, M.hdr_src.data(), M.q_src.data(), M.chk.data(), vk_cap, base, off, w, &bad);
    if (malformed) *malformed = bad ? 1 : 0;
    {
        std::vector<u64> v(M.d.num_public_inputs);
        for (u32 j = 0; j < M.d.num_public_inputs; j++) {
            v[j] = wire_gather64(base, off + M.d.pi_off + 8 * (size_t)j);
            if (!is_canonical(v[j]) && malformed) *malformed = 1;   // assign_value range check, arithmetic_chip.rs:256-268
        }
        // PublicInputsHasherChip is Poseidon-Goldilocks whatever the proof's Merkle hasher is
        // (chip/public_inputs_hasher_chip.rs:35-58, InnerHasher = PoseidonHash plonky2_config.rs:74)
        if (pi_hash) hash_no_pad(SV_HASH_POSEIDON_GOLDILOCKS, v.data(), v.size(), pi_hash);
        if (pis && !v.empty()) memcpy(pis, v.data(), v.size() * 8);
    }
}

}  // namespace svb

using namespace svb;

extern "C" int sv_fri_shape_from_common(const sv_plonk_common* c, uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                                        uint32_t num_query_rounds, uint32_t proof_of_work_bits, uint32_t num_steps,
                                        const uint32_t* reduction_arity_bits, uint32_t hiding, uint32_t hash_kind, sv_fri_shape* out) {
    if (!c || !out) return -1;
    if (num_steps > degree_bits || num_steps > SV_MAX_STEPS) return -2;
    sv_fri_shape s;
    memset(&s, 0, sizeof s);
    s.degree_bits = degree_bits;
    s.rate_bits = rate_bits;
    s.cap_height = cap_height;
    s.num_query_rounds = num_query_rounds;
    s.proof_of_work_bits = proof_of_work_bits;
    s.num_steps = num_steps;
    uint32_t total = 0;
    for (uint32_t i = 0; i < num_steps; i++) {
        s.reduction_arity_bits[i] = reduction_arity_bits ? reduction_arity_bits[i] : 1u;
        total += s.reduction_arity_bits[i];
    }
    if (total > degree_bits) return -2;
    s.final_poly_len = 1u << (degree_bits - total);
    s.hiding = hiding ? 1 : 0;
    s.oracle_num_polys[0] = c->num_constants + c->num_routed_wires;
    s.oracle_num_polys[1] = c->num_wires;
    s.oracle_num_polys[2] = c->num_challenges * (1 + c->num_partial_products);
    s.oracle_num_polys[3] = c->num_challenges * c->quotient_degree_factor;
    s.oracle_blinding[0] = 0;  // PlonkOracle::CONSTANTS_SIGMAS .. QUOTIENT, types/common_data.rs:104-122
    s.oracle_blinding[1] = s.oracle_blinding[2] = s.oracle_blinding[3] = 1;
    s.num_zs = c->num_challenges;
    s.hash_kind = hash_kind;
    sv_fri_layout L;
    if (make_layout(s, L)) return -2;
    *out = s;
    return 0;
}

extern "C" size_t sv_wire_proof_bytes(const sv_fri_shape* shape, const sv_plonk_common* common) {
    WireMap M;
    if (!shape || !common || make_wire_map(*shape, *common, M)) return 0;
    return M.d.proof_bytes;
}

extern "C" int sv_public_inputs_hash(const uint64_t* public_inputs, size_t n, uint64_t out[4]) {
    if (!out || (n && !public_inputs)) return -1;
    hash_no_pad(SV_HASH_POSEIDON_GOLDILOCKS, public_inputs, n, out);
    return 0;
}

extern "C" int sv_wire_pack(const sv_fri_shape* shape, const sv_plonk_common* common, const uint64_t* rec,
                            const uint64_t* public_inputs, uint8_t* out) {
    WireMap M;
    if (!shape || !common || !rec || !out) return -1;
    if (int rc = make_wire_map(*shape, *common, M)) return rc;
    const WireDims& d = M.d;
    if (d.num_public_inputs && !public_inputs) return -1;
    memset(out, 0, d.proof_bytes);
    for (u32 w = 0; w < d.header_words; w++) {
        u32 code = M.hdr_src[w];
        if (code != WIRE_ZERO && !(code & WIRE_VK_FLAG)) memcpy(out + code, rec + w, 8);
    }
    for (u32 q = 0; q < d.num_queries; q++) {
        uint8_t* qb = out + d.query_base + (size_t)q * d.query_bytes;
        const u64* qr = rec + d.header_words + (size_t)q * d.query_words;
        for (u32 r = 0; r < d.query_words; r++)
            if (M.q_src[r] != WIRE_ZERO) memcpy(qb + M.q_src[r], qr + r, 8);
        for (u32 j = 0; j < d.n_chk; j++) qb[M.chk[j] >> 8] = (uint8_t)(M.chk[j] & 0xFFu);
    }
    if (d.num_public_inputs) memcpy(out + d.pi_off, public_inputs, 8 * (size_t)d.num_public_inputs);
    return 0;
}

extern "C" int sv_wire_unpack_batch(const sv_fri_shape* shape, const sv_plonk_common* common, const uint64_t* vk_cap,
                                    const uint8_t* blob, size_t stride, size_t n, uint64_t* records_out, uint64_t* pi_hashes_out,
                                    uint64_t* public_inputs_out, uint8_t* malformed_out, int nthreads) {
    WireMap M;
    if (!shape || !common || !vk_cap || !records_out || (n && !blob)) return -1;
    if (int rc = make_wire_map(*shape, *common, M)) return rc;
    const WireDims& d = M.d;
    if (n > 1 && stride < d.proof_bytes) return -4;
    if (nthreads < 1) nthreads = 1;
    parallel_for(n, nthreads, [&](size_t b, size_t e) {
        std::vector<u64> tmp((d.proof_bytes + 7) / 8 + 1);
        for (size_t i = b; i < e; i++) {
            const uint8_t* p = blob + i * stride;
            const u64* base;
            size_t off;
            if (i == 0 || i + 1 == n) {
                // the aligned words around the first / last proof may reach outside the caller's buffer
                tmp.back() = 0;
                memcpy(tmp.data(), p, d.proof_bytes);
                base = tmp.data();
                off = 0;
            } else {
                base = reinterpret_cast<const u64*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
                off = reinterpret_cast<uintptr_t>(p) & 7;
            }
            unpack_one(M, vk_cap, base, off, records_out + i * (size_t)d.record_words, pi_hashes_out ? pi_hashes_out + 4 * i : nullptr,
                       public_inputs_out ? public_inputs_out + i * (size_t)d.num_public_inputs : nullptr,
                       malformed_out ? malformed_out + i : nullptr);
        }
    });
    return 0;
}
