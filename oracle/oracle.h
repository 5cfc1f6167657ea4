/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path never does.
 *
 * Public surface of the CPU restatement of the reference's FRI query-phase verifier
 * (chip/fri_chip.rs, chip/merkle_proof_chip.rs, chip/hasher_chip.rs) with the Poseidon-Goldilocks
 * permutation (constants + fast algorithm: chip/plonk/gates/poseidon.rs:26-322,504-589,634-686).
 *
 * Parity status: the Poseidon permutation, the field and the sponge are PINNED by the known-answer
 * vectors of SURVEY.md section 8c (the first two equal upstream plonky2's published test_vectors12).
 * The end-to-end FRI accept/reject is NOT pinned against a real plonky2 proof: the reference is Rust
 * (no cargo/rustc in this image, plonky2 and halo2 sources not on disk, no serialized proofs or
 * golden vectors anywhere in the reference) -- "parity unpinned" at that level; the restatement is
 * authoritative by construction-from-citation.
 */
#ifndef ORC_ORACLE_H
#define ORC_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* == FriParams (types/common_data.rs:43-54) + FriInstanceInfo (types/fri.rs:50-72) == */
typedef struct {
    uint32_t degree_bits;        /* FriParams.degree_bits */
    uint32_t rate_bits;          /* FriConfig.rate_bits */
    uint32_t cap_height;         /* FriConfig.cap_height */
    uint32_t num_query_rounds;   /* FriConfig.num_query_rounds */
    uint32_t proof_of_work_bits; /* FriConfig.proof_of_work_bits */
    uint32_t num_steps;          /* reduction_arity_bits.len() */
    uint32_t final_poly_len;     /* number of Fp2 coefficients of final_poly */
    uint32_t hiding;             /* FriParams.hiding */
    uint32_t oracle_num_polys[4];/* FriOracleInfo.num_polys, order: constants_sigmas, wires, zs_pp, quotient */
    uint32_t oracle_blinding[4]; /* FriOracleInfo.blinding */
    uint32_t num_zs;             /* batch 1 (point g*zeta) = polys [0,num_zs) of oracle 2 (common_data.rs:192-194) */
    uint32_t hash_kind;          /* 0 = Poseidon-Goldilocks, 1 = Poseidon-BN254 wrapped (family B) */
    uint32_t reduction_arity_bits[32]; /* FriParams.reduction_arity_bits (types/common_data.rs:47); fri_chip.rs:211 implements
                                    * arity 2 only, larger arities follow plonky2's compute_evaluation (see oracle.c) */
} orc_shape;

/* Word offsets (u64) of the flat per-proof record; same format as include/stark_verifier_b200.h
 * but computed by independent code so the two can be cross-checked. */
typedef struct {
    uint32_t ncap, lde_bits, n0, n1;
    uint32_t off_init_caps, off_step_caps, off_open0, off_open1, off_final_poly, off_pow_witness;
    uint32_t off_alpha, off_betas, off_pow_response, off_indices, off_zeta, off_zeta_next;
    uint32_t header_words;
    uint32_t leaf_len[4];
    uint32_t q_off_init_evals[4], q_off_init_sibs[4], init_depth;
    uint32_t q_off_step_evals[32], q_off_step_sibs[32], step_depth[32];
    uint32_t query_words, record_words;
} orc_layout;

int orc_make_layout(const orc_shape *s, orc_layout *L);

/* field restatement (orc_field.h), exported for the KAT tests */
int orc_tables_canonical(void);
uint64_t orc_f_mul(uint64_t a, uint64_t b);
uint64_t orc_f_pow(uint64_t a, uint64_t e);
uint64_t orc_f_inv(uint64_t a);
uint64_t orc_f_red128(uint64_t lo, uint64_t hi, int slow);
void orc_f2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);
void orc_f2_inv(const uint64_t a[2], uint64_t out[2]);

/* Poseidon-Goldilocks permutation: fast form (poseidon.rs:634-686) and naive 30-round form. */
void orc_poseidon(uint64_t st[12]);
void orc_poseidon_naive(uint64_t st[12]);
void orc_poseidon_batch(uint64_t *states, size_t n);
/* Hash family B (bn245_poseidon): raw Fr permutation on canonical integers (state[5][4] little-endian
 * u64 limbs) and the Goldilocks-wrapped width-12 permutation (plonky2_config.rs:38-51). */
void orc_poseidon_b_fr(uint64_t state[20]);
void orc_poseidon_b(uint64_t st[12]);
/* Select the permutation used by orc_hash_no_pad / orc_two_to_one / orc_merkle_verify* on the calling
 * thread (0 = Poseidon-Goldilocks, 1 = B).  orc_fri_verify / orc_fri_challenges set it from
 * orc_shape.hash_kind themselves. */
void orc_set_hash_kind(int kind);
/* hash_n_to_m_no_pad, overwrite-mode sponge (hasher_chip.rs:122-148) -> 4 outputs */
void orc_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]);
/* two_to_one == HasherChip::permute on a fresh zero state (hasher_chip.rs:150-171, merkle_proof_chip.rs:58-71) */
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
/* merkle_proof_chip.rs:39-87.  index = the leaf_index_bits as an integer (LSB first); returns 1 iff ok */
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, uint64_t index, const uint64_t *siblings,
                      size_t depth, const uint64_t *cap, size_t cap_index);
void orc_merkle_verify_batch(const uint64_t *records, size_t n, size_t leaf_len, size_t depth,
                             const uint64_t *indices, const uint64_t *caps, size_t cap_height, uint8_t *ok);

/* fail codes (0 == accept) */
enum {
    ORC_OK = 0, ORC_FAIL_POW = 1, ORC_FAIL_NONCANONICAL = 2, ORC_FAIL_INIT_MERKLE = 3,
    ORC_FAIL_ZERO_DENOM = 4, ORC_FAIL_STEP_EVAL = 5, ORC_FAIL_STEP_MERKLE = 6, ORC_FAIL_FINAL = 7
};
/* FriVerifierChip::verify_fri_proof (fri_chip.rs:329-362) on one flat record.  returns 1 iff accepted;
 * *fail (optional) = first failing check, *fail_query = query round it occurred in (or -1). */
int orc_fri_verify(const orc_shape *s, const uint64_t *record, int *fail, int *fail_query);
/* batch, packed bitmap (bit i of word i/32 = proof i accepted); nthreads pthreads */
void orc_fri_verify_batch(const orc_shape *s, const uint64_t *records, size_t n, uint32_t *bitmap, int nthreads);

/* Fiat-Shamir (plonk_verifier_chip.rs:55-154, transcript_chip.rs, hasher_chip.rs:51-89).
 * Reads caps/openings/final_poly/pow_witness from the record header and WRITES
 * zeta, zeta_next (= g*zeta), fri_alpha, fri_betas, fri_pow_response, fri_query_indices into it.
 * circuit_digest[4], pi_hash[4]: observed first (plonk_verifier_chip.rs:65-71). */
void orc_fri_challenges(const orc_shape *s, uint64_t *record, const uint64_t circuit_digest[4],
                        const uint64_t pi_hash[4], uint32_t num_challenges);

/* == wire format (wire.c): plonky2's ProofWithPublicInputs bytes <-> flat record == */
typedef struct {
    uint32_t num_constants, num_routed_wires, num_wires, num_challenges, num_partial_products,
        quotient_degree_factor, num_public_inputs; /* CommonData / CircuitConfig, types/common_data.rs:23-40,68-96 */
} orc_common;
size_t orc_wire_proof_bytes(const orc_shape *s, const orc_common *c);
int orc_wire_read_proof(const orc_shape *s, const orc_common *c, const uint64_t *vk_constants_sigmas_cap,
                        const uint8_t *bytes, size_t len, uint64_t *record, uint64_t *public_inputs, uint64_t pi_hash[4]);
int orc_wire_write_proof(const orc_shape *s, const orc_common *c, const uint64_t *record, const uint64_t *public_inputs,
                         uint8_t *out);

/* == plonk-level checks (plonk.c): the vanishing-polynomial identity at zeta == */
typedef struct { uint32_t kind, param, param2, param3, selector_index; } orc_plonk_gate; /* kind: 0 noop, 1 constant, 2 public input, 3 arithmetic, 4 arithmetic ext, 5 mul ext, 6 base sum (base 2), 7 reducing, 8 reducing ext,
 * 9 random access (bits, copies, extra constants), 10 poseidon mds, 11 poseidon */
typedef struct {
    orc_common common;
    uint32_t degree_bits, num_gate_constraints, num_selectors;
    uint32_t group_lo[8], group_hi[8]; /* SelectorsInfo.groups */
    uint32_t num_gates;
    orc_plonk_gate gates[32];          /* CommonData.gates order */
    uint64_t k_is[128];
} orc_plonk_circuit;
/* open0 / open1: FriOpenings batches (types/assigned.rs:26-40); chal = betas | gammas | alphas; 1 = identity holds */
int orc_plonk_check(const orc_plonk_circuit *C, const uint64_t *open0, const uint64_t *open1, const uint64_t pi_hash[4],
                    const uint64_t *chal, const uint64_t zeta[2]);
/* the plonk challenges orc_fri_challenges squeezes on its way to zeta: out = betas | gammas | alphas */
void orc_plonk_challenges(const orc_shape *s, const uint64_t *record, const uint64_t circuit_digest[4],
                          const uint64_t pi_hash[4], uint32_t num_challenges, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
