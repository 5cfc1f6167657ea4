/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Cursor-style restatement of plonky2's proof (de)serialiser for the fields the FRI path reads, written
 * independently of the product's table-driven packer (stark-verifier_b200/csrc/wire.hpp) so that the two can be
 * cross-checked.  Follows util/serialization.rs of the plonky2 revision the reference pins (Cargo.lock:
 * DoHoonKim8/plonky2#72229c47, not on disk -- restated from its published source, pre-lookup format):
 *   write_proof_with_public_inputs = write_proof, then write_field_vec(public_inputs)
 *   write_proof        = write_merkle_cap(wires_cap), (plonk_zs_partial_products_cap), (quotient_polys_cap),
 *                        write_opening_set, write_fri_proof
 *   write_opening_set  = write_field_ext_vec of constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next,
 *                        partial_products, quotient_polys          (lengths come from CommonCircuitData on read)
 *   write_fri_proof    = commit_phase_merkle_caps, query rounds, final_poly coefficients, pow_witness
 *   query round        = 4 x (write_field_vec(evals), write_merkle_proof), then per step
 *                        (write_field_ext_vec(evals), write_merkle_proof)
 *   write_merkle_proof = u8 length, then the sibling hashes;  write_field = canonical u64, little-endian;
 *   hash = 4 fields; extension element = 2 fields.
 * The destination is the flat record of oracle.h (orc_layout), i.e. the reference's ProofValues::from
 * (types/proof.rs:389-403) and VerificationKeyValues::from (types/verification_key.rs:14-24) for those fields:
 * batch 0 of the FRI openings is [constants, sigmas, wires, zs, partial_products, quotient], batch 1 is zs_next
 * (types/assigned.rs:26-40).  Parity unpinned: the reference ships no serialised proof to check against. */
#include "oracle.h"
#include <string.h>

typedef struct {
    const uint8_t *p;
    size_t len, pos;
    int short_read;
} rd_t;

static uint64_t rd_u64(rd_t *r) {
    if (r->pos + 8 > r->len) { r->short_read = 1; r->pos = r->len; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)r->p[r->pos + i] << (8 * i);
    r->pos += 8;
    return v;
}
static unsigned rd_u8(rd_t *r) {
    if (r->pos + 1 > r->len) { r->short_read = 1; return 0; }
    return r->p[r->pos++];
}
static void rd_fields(rd_t *r, uint64_t *dst, size_t n) { for (size_t i = 0; i < n; i++) dst[i] = rd_u64(r); }

typedef struct { uint8_t *p; size_t pos; } wr_t;
static void wr_u64(wr_t *w, uint64_t v) { for (int i = 0; i < 8; i++) w->p[w->pos++] = (uint8_t)(v >> (8 * i)); }
static void wr_fields(wr_t *w, const uint64_t *src, size_t n) { for (size_t i = 0; i < n; i++) wr_u64(w, src[i]); }

static const uint64_t P = 0xFFFFFFFF00000001ull;

size_t orc_wire_proof_bytes(const orc_shape *s, const orc_common *c) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return 0;
    size_t n = 0;
    n += 3 * (size_t)L.ncap * 32;
    n += 16 * (size_t)(c->num_constants + c->num_routed_wires + c->num_wires + 2 * c->num_challenges +
                       c->num_challenges * c->num_partial_products + c->num_challenges * c->quotient_degree_factor);
    n += (size_t)s->num_steps * L.ncap * 32;
    size_t q = 0;
    for (int k = 0; k < 4; k++) q += 8 * (size_t)L.leaf_len[k] + 1 + 32 * (size_t)L.init_depth;
    for (uint32_t i = 0; i < s->num_steps; i++) q += ((size_t)16 << s->reduction_arity_bits[i]) + 1 + 32 * (size_t)L.step_depth[i];
    n += q * s->num_query_rounds;
    n += 16 * (size_t)s->final_poly_len + 8;
    n += 8 * (size_t)c->num_public_inputs;
    return n;
}

/* returns < 0: the bytes cannot be framed (length); 0: parsed; 1: parsed but malformed (a Merkle-proof length that
 * disagrees with the FRI parameters, or a public input >= p) */
int orc_wire_read_proof(const orc_shape *s, const orc_common *c, const uint64_t *vk_constants_sigmas_cap,
                        const uint8_t *bytes, size_t len, uint64_t *rec, uint64_t *public_inputs, uint64_t pi_hash[4]) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return -1;
    if (len != orc_wire_proof_bytes(s, c)) return -2;
    memset(rec, 0, (size_t)L.record_words * 8);
    rd_t r = {bytes, len, 0, 0};
    int malformed = 0;
    const size_t capw = (size_t)L.ncap * 4;
    /* VerificationKeyValues.constants_sigmas_cap is initial_merkle_caps[0] (plonk_verifier_chip.rs:212-217) */
    memcpy(rec + L.off_init_caps, vk_constants_sigmas_cap, capw * 8);
    rd_fields(&r, rec + L.off_init_caps + 1 * capw, capw); /* wires_cap */
    rd_fields(&r, rec + L.off_init_caps + 2 * capw, capw); /* plonk_zs_partial_products_cap */
    rd_fields(&r, rec + L.off_init_caps + 3 * capw, capw); /* quotient_polys_cap */
    /* read_opening_set */
    uint64_t *o0 = rec + L.off_open0, *o1 = rec + L.off_open1;
    rd_fields(&r, o0, 2 * (size_t)c->num_constants);        o0 += 2 * c->num_constants;
    rd_fields(&r, o0, 2 * (size_t)c->num_routed_wires);     o0 += 2 * c->num_routed_wires;
    rd_fields(&r, o0, 2 * (size_t)c->num_wires);            o0 += 2 * c->num_wires;
    rd_fields(&r, o0, 2 * (size_t)c->num_challenges);       o0 += 2 * c->num_challenges;       /* plonk_zs */
    rd_fields(&r, o1, 2 * (size_t)c->num_challenges);                                           /* plonk_zs_next */
    rd_fields(&r, o0, 2 * (size_t)c->num_challenges * c->num_partial_products);   o0 += 2 * c->num_challenges * c->num_partial_products;
    rd_fields(&r, o0, 2 * (size_t)c->num_challenges * c->quotient_degree_factor); o0 += 2 * c->num_challenges * c->quotient_degree_factor;
    if ((size_t)(o0 - (rec + L.off_open0)) != 2 * (size_t)L.n0 || c->num_challenges != L.n1) return -3;
    /* read_fri_proof */
    rd_fields(&r, rec + L.off_step_caps, (size_t)s->num_steps * capw);
    for (uint32_t q = 0; q < s->num_query_rounds; q++) {
        uint64_t *qp = rec + L.header_words + (size_t)q * L.query_words;
        for (int k = 0; k < 4; k++) { /* read_fri_initial_proof */
            rd_fields(&r, qp + L.q_off_init_evals[k], L.leaf_len[k]);
            if (rd_u8(&r) != L.init_depth) malformed = 1;
            rd_fields(&r, qp + L.q_off_init_sibs[k], 4 * (size_t)L.init_depth);
        }
        for (uint32_t i = 0; i < s->num_steps; i++) { /* read_fri_query_step: 2^arity_bits extension evals */
            rd_fields(&r, qp + L.q_off_step_evals[i], (size_t)2 << s->reduction_arity_bits[i]);
            if (rd_u8(&r) != L.step_depth[i]) malformed = 1;
            rd_fields(&r, qp + L.q_off_step_sibs[i], 4 * (size_t)L.step_depth[i]);
        }
    }
    rd_fields(&r, rec + L.off_final_poly, 2 * (size_t)s->final_poly_len);
    rec[L.off_pow_witness] = rd_u64(&r);
    /* public inputs and their hash (PlonkVerifierChip::get_public_inputs_hash, plonk_verifier_chip.rs:41-53:
     * Poseidon-Goldilocks whatever the proof's hasher) */
    uint64_t pis[c->num_public_inputs ? c->num_public_inputs : 1];
    rd_fields(&r, pis, c->num_public_inputs);
    for (uint32_t i = 0; i < c->num_public_inputs; i++)
        if (pis[i] >= P) malformed = 1;
    if (r.short_read || r.pos != len) return -4;
    if (public_inputs) memcpy(public_inputs, pis, 8 * (size_t)c->num_public_inputs);
    if (pi_hash) {
        orc_set_hash_kind(0);
        orc_hash_no_pad(pis, c->num_public_inputs, pi_hash);
        orc_set_hash_kind((int)s->hash_kind);
    }
    return malformed;
}

/* the inverse: record + public inputs -> bytes (ProofWithPublicInputs::to_bytes order) */
int orc_wire_write_proof(const orc_shape *s, const orc_common *c, const uint64_t *rec, const uint64_t *public_inputs,
                         uint8_t *out) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return -1;
    wr_t w = {out, 0};
    const size_t capw = (size_t)L.ncap * 4;
    wr_fields(&w, rec + L.off_init_caps + capw, 3 * capw);
    const uint64_t *o0 = rec + L.off_open0;
    size_t head = 2 * (size_t)(c->num_constants + c->num_routed_wires + c->num_wires + c->num_challenges);
    wr_fields(&w, o0, head);
    wr_fields(&w, rec + L.off_open1, 2 * (size_t)c->num_challenges);
    wr_fields(&w, o0 + head, 2 * (size_t)L.n0 - head);
    wr_fields(&w, rec + L.off_step_caps, (size_t)s->num_steps * capw);
    for (uint32_t q = 0; q < s->num_query_rounds; q++) {
        const uint64_t *qp = rec + L.header_words + (size_t)q * L.query_words;
        for (int k = 0; k < 4; k++) {
            wr_fields(&w, qp + L.q_off_init_evals[k], L.leaf_len[k]);
            w.p[w.pos++] = (uint8_t)L.init_depth;
            wr_fields(&w, qp + L.q_off_init_sibs[k], 4 * (size_t)L.init_depth);
        }
        for (uint32_t i = 0; i < s->num_steps; i++) {
            wr_fields(&w, qp + L.q_off_step_evals[i], (size_t)2 << s->reduction_arity_bits[i]);
            w.p[w.pos++] = (uint8_t)L.step_depth[i];
            wr_fields(&w, qp + L.q_off_step_sibs[i], 4 * (size_t)L.step_depth[i]);
        }
    }
    wr_fields(&w, rec + L.off_final_poly, 2 * (size_t)s->final_poly_len);
    wr_u64(&w, rec[L.off_pow_witness]);
    wr_fields(&w, public_inputs, c->num_public_inputs);
    return w.pos == orc_wire_proof_bytes(s, c) ? 0 : -2;
}
