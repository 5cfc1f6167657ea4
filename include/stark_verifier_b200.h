/* stark_verifier_b200.h -- C ABI of the B200-native batch FRI query-phase verifier.
 *
 * The reference (DoHoonKim8/stark-verifier, crate `semaphore_aggregation`) has NO FFI today; this
 * header is the extern "C" seam a maintainer would bind from Rust.  Each entry
 * point names the reference interface it replaces (paths relative to src/plonky2_verifier/).
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every input/output buffer; the library owns
 *    only `sv_ctx` (CUDA stream, staging buffers, constant tables);
 *  - return value: 0 = ok, < 0 = error (bad shape / CUDA failure); `sv_last_error` gives text;
 *  - an INVALID PROOF IS DATA, NOT AN ERROR: bit i of the accept bitmap says whether proof i was
 *    accepted (reference: panic / failed MockProver, verifier_api.rs:50-51);
 *  - every field element is a canonical little-endian u64 < p = 2^64 - 2^32 + 1; a non-canonical
 *    word anywhere in a proof rejects that proof (reference: range check in assign_value,
 *    native_chip/arithmetic_chip.rs:256-268);
 *  - `mem` says where the data buffers live: SV_MEM_HOST (the call copies H2D in chunks overlapped
 *    with the kernels, copies the result back and returns when it is in host memory) or
 *    SV_MEM_DEVICE (pointers are device pointers on the ctx device; the work is enqueued on the
 *    ctx stream and the call returns without synchronising; the calls of one ctx share its scratch buffers, so the library
 *    orders a call behind the previous one even when sv_ctx_set_stream moved the ctx to another stream in between);
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef STARK_VERIFIER_B200_H
#define STARK_VERIFIER_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SV_MEM_HOST 0
#define SV_MEM_DEVICE 1

#define SV_HASH_POSEIDON_GOLDILOCKS 0 /* plonky2 PoseidonHash; constants chip/plonk/gates/poseidon.rs:26-322 */
#define SV_HASH_POSEIDON_BN254 1      /* Poseidon over BN254 Fr wrapped around 12 Goldilocks limbs: bn245_poseidon/plonky2_config.rs:38-66,
                                         native.rs:16-77, constants.rs (the Hasher of Bn254PoseidonGoldilocksConfig) */

#define SV_MAX_STEPS 32
#define SV_MAX_ARITY_BITS 4 /* a reduction step folds at most 16 evaluations (plonky2's strategies stop at arity 16) */

/* == FriParams + FriConfig (types/common_data.rs:10-54) and the part of FriInstanceInfo
 *    (types/fri.rs:50-72, types/common_data.rs:153-221) the FRI verifier reads. == */
typedef struct sv_fri_shape {
    uint32_t degree_bits;         /* FriParams.degree_bits */
    uint32_t rate_bits;           /* FriConfig.rate_bits */
    uint32_t cap_height;          /* FriConfig.cap_height */
    uint32_t num_query_rounds;    /* FriConfig.num_query_rounds */
    uint32_t proof_of_work_bits;  /* FriConfig.proof_of_work_bits */
    uint32_t num_steps;           /* FriParams.reduction_arity_bits.len() */
    uint32_t final_poly_len;      /* number of Fp2 coefficients in FriProofValues.final_poly
                                     = 2^(degree_bits - sum of reduction_arity_bits) for a plonky2 proof */
    uint32_t hiding;              /* FriParams.hiding */
    uint32_t oracle_num_polys[4]; /* FriOracleInfo.num_polys: constants_sigmas, wires, zs_partial_products, quotient */
    uint32_t oracle_blinding[4];  /* FriOracleInfo.blinding (salted leaf = +4 limbs when hiding) */
    uint32_t num_zs;              /* batch 1 (opened at g*zeta) = polynomials [0, num_zs) of oracle 2 */
    uint32_t hash_kind;           /* SV_HASH_* */
    uint32_t reduction_arity_bits[SV_MAX_STEPS]; /* FriParams.reduction_arity_bits[0 .. num_steps): 1 .. SV_MAX_ARITY_BITS
                                     each.  The reference's in-circuit verifier folds arity 2 only (chip/fri_chip.rs:211);
                                     its own demo proves with ConstantArityBits(3, 5) (plonky2_semaphore/access_set.rs:124)
                                     and checks those proofs with plonky2's native verifier, whose 2^k-ary fold
                                     (compute_evaluation: interpolate the coset, evaluate at beta) is what runs here. */
} sv_fri_shape;

/* Flat per-proof record: word (u64) offsets.  Every segment starts on a 4-word (32-byte) boundary
 * so that digests can be fetched with aligned 128/256-bit loads.
 *
 *   header (shared by all query rounds of the proof):
 *     init_caps    4 x ncap x 4     MerkleCapValues of the 4 initial oracles, in FriInstanceInfo order
 *                                   (plonk_verifier_chip.rs:212-217)
 *     step_caps    num_steps x ncap x 4   FriProofValues.commit_phase_merkle_cap_values
 *     open0        n0 x 2           FriOpenings.batches[0] (order: types/assigned.rs:26-37)
 *     open1        n1 x 2           FriOpenings.batches[1] (plonk_zs_next)
 *     final_poly   final_poly_len x 2
 *     pow_witness  1
 *     alpha 2, betas num_steps x 2, pow_response 1, indices num_query_rounds   FriChallenges
 *     zeta 2, zeta_next 2           FriBatchInfo.point of the two batches
 *   then num_query_rounds x query block (FriQueryRoundValues, types/proof.rs:217):
 *     for k in 0..4: evals[leaf_len[k]], siblings[init_depth x 4]    (FriInitialTreeProofValues)
 *     for i in steps: evals[2^arity_bits[i] x 2], siblings[step_depth[i] x 4]      (FriQueryStepValues; the
 *                                   evals are the step tree's leaf: more than 4 words are hashed, 4 are their own digest)
 */
typedef struct sv_fri_layout {
    uint32_t ncap, lde_bits, n0, n1;
    uint32_t off_init_caps, off_step_caps, off_open0, off_open1, off_final_poly, off_pow_witness;
    uint32_t off_alpha, off_betas, off_pow_response, off_indices, off_zeta, off_zeta_next;
    uint32_t header_words;
    uint32_t leaf_len[4];
    uint32_t q_off_init_evals[4], q_off_init_sibs[4], init_depth;
    uint32_t q_off_step_evals[SV_MAX_STEPS], q_off_step_sibs[SV_MAX_STEPS], step_depth[SV_MAX_STEPS];
    uint32_t step_arity_bits[SV_MAX_STEPS]; /* copy of the shape's reduction_arity_bits */
    uint32_t step_index_shift[SV_MAX_STEPS]; /* leaf index of step tree i = x_index >> step_index_shift[i] (sum of arity bits up to i) */
    uint32_t query_words, record_words;
    /* algorithmic HBM read bytes: unpadded per-query and per-proof-shared payload */
    uint32_t algo_bytes_per_query, algo_bytes_shared;
    uint32_t perms_per_query; /* Poseidon permutations per (proof x query) */
} sv_fri_layout;

/* first-failure codes, ordered like the checks of check_consistency (chip/fri_chip.rs:228-327) */
enum {
    SV_OK = 0,
    SV_FAIL_POW = 1,          /* fri_verify_proof_of_work, fri_chip.rs:364-376 */
    SV_FAIL_NONCANONICAL = 2, /* a word >= p */
    SV_FAIL_INIT_MERKLE = 3,  /* verify_initial_merkle_proof, fri_chip.rs:85-110 */
    SV_FAIL_ZERO_DENOM = 4,   /* div_extension on zero, goldilocks_extension_chip.rs:72-101 */
    SV_FAIL_STEP_EVAL = 5,    /* evals[x_index_within_coset] != prev_eval, fri_chip.rs:285-292 */
    SV_FAIL_STEP_MERKLE = 6,  /* step Merkle proof, fri_chip.rs:303-311 */
    SV_FAIL_FINAL = 7         /* final_poly(x) != prev_eval, fri_chip.rs:317-325 */
};
/* first_fail[i] = 0 if proof i is accepted, else (query_round << 8) | code of the FIRST check that
 * fails in the reference's order (query rounds in order; inside a round: non-canonical word, the 4
 * initial Merkle proofs, DEEP-quotient division, then per step: eval consistency, fold division,
 * step Merkle proof; finally the final polynomial).  Per-proof failures (PoW, header word >= p) use
 * query_round = 0. */

typedef struct sv_ctx sv_ctx;

/* --- context --------------------------------------------------------------------------------- */
int sv_ctx_create(int device, sv_ctx** out);
void sv_ctx_destroy(sv_ctx* ctx);
const char* sv_last_error(const sv_ctx* ctx); /* ctx may be NULL: last error of a failed create */
/* Use an existing CUDA stream (cudaStream_t cast to void*) for SV_MEM_DEVICE work; NULL restores
 * the context's own stream. */
int sv_ctx_set_stream(sv_ctx* ctx, void* cuda_stream);
int sv_ctx_synchronize(sv_ctx* ctx);
/* number of kernel launches issued through this context so far */
uint64_t sv_ctx_launch_count(const sv_ctx* ctx);
/* CUDA-event timing of the dominant kernel of each call (fri_query_kernel, merkle_verify_kernel,
 * poseidon_permute_kernel), recorded on the stream the kernel is launched on.  kernel_time_ms
 * synchronises, returns the sum and the number of launches since the last call, and resets. */
int sv_ctx_kernel_timing(sv_ctx* ctx, int enable);
int sv_ctx_kernel_time_ms(sv_ctx* ctx, double* total_ms, uint64_t* n_launches);
/* pinned host memory for the SV_MEM_HOST path (pageable memory works too, slower) */
int sv_host_alloc(size_t bytes, void** out);
int sv_host_free(void* p);

/* --- layout ---------------------------------------------------------------------------------- */
int sv_fri_layout_make(const sv_fri_shape* shape, sv_fri_layout* out);

/* --- hot path -------------------------------------------------------------------------------- */
/* n independent width-12 permutations, in[12n] -> out[12n] (in == out allowed).
 * Replaces: PlonkyPermutation::permute / HasherChip::permutation (chip/hasher_chip.rs:101-105). */
int sv_poseidon_permute_batch(sv_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, int hash_kind, int mem);
/* The same permutation (Poseidon-Goldilocks only) on the lane-cooperative mapping the device-side transcript uses -- 16 lanes
 * per state, built for latency (HasherChip's duplex sponge, chip/hasher_chip.rs:51-120, is a chain of dependent permutations).
 * Same results as sv_poseidon_permute_batch; in and out must not overlap for SV_MEM_DEVICE.  Exposed for the parity tests. */
int sv_poseidon_permute_batch_coop(sv_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, int mem);

/* out[i] = a[i] * b[i] + c[i] mod p (canonical); a, b, c are arbitrary u64 words (reduced mod p).
 * Replaces: GoldilocksChip::mul_add, the gate r = a*b + c (chip/goldilocks_chip.rs:175-195,
 * native_chip/arithmetic_chip.rs:98-107) -- the arithmetic every kernel of this library is built on. */
int sv_goldilocks_mul_add_batch(sv_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* out,
                                size_t n, int mem);

/* n independent Merkle paths against one cap.  paths: n records of (leaf_len + 4*depth) words
 * (leaf, then siblings bottom-up); indices[n] leaf indices (bit l of the index selects the side at
 * level l; the bits above `depth` select the cap entry); caps: 2^cap_height x 4 words; ok[n] bytes.
 * Replaces: MerkleProofChip::verify_merkle_proof_to_cap_with_cap_index (chip/merkle_proof_chip.rs:39-87). */
int sv_merkle_verify_batch(sv_ctx* ctx, uint32_t leaf_len, uint32_t depth, uint32_t cap_height, int hash_kind,
                           const uint64_t* paths, const uint64_t* indices, const uint64_t* caps, uint8_t* ok,
                           size_t n, int mem);

/* Build the Merkle tree of n_leaves = 2^k leaf rows of leaf_len words (row-major, contiguous).
 * layers_out receives the digest layers bottom-up: leaf digests (n_leaves x 4 words; a row of <= 4 words is its own
 * digest, zero-padded), then every level of two_to_one parents down to the cap layer (2^cap_height x 4 words):
 * 4 * (2 * n_leaves - 2^cap_height) words in total.  The sibling of node j at a level is node j ^ 1 of that layer.
 * Replaces: plonky2 MerkleTree::new (call sites plonky2_semaphore/access_set.rs:25, circuit.rs:91) -- the prover
 * side of sv_merkle_verify_batch. */
int sv_merkle_tree_build(sv_ctx* ctx, int hash_kind, uint32_t leaf_len, const uint64_t* leaves, size_t n_leaves,
                         uint32_t cap_height, uint64_t* layers_out, int mem);

/* Batch FRI query-phase verification: n_proofs flat records (sv_fri_layout.record_words each).
 * accept_bitmap: ceil(n_proofs/32) u32 words, bit (i & 31) of word i/32 = proof i accepted.
 * first_fail: optional (NULL to skip), n_proofs u32 (see above).
 * Replaces: FriVerifierChip::verify_fri_proof (chip/fri_chip.rs:329-362) called once per proof from
 * PlonkVerifierChip::verify_proof_with_challenges (chip/plonk/plonk_verifier_chip.rs:212-240). */
int sv_fri_verify_batch(sv_ctx* ctx, const sv_fri_shape* shape, size_t n_proofs, const uint64_t* records,
                        uint32_t* accept_bitmap, uint32_t* first_fail, int mem);

/* Device-side Fiat-Shamir: the transcript of sv_fri_challenges for n_proofs records at once,
 * one GPU thread per proof; rewrites zeta, zeta_next, alpha, betas, pow_response and indices of every
 * record header in place.  circuit_digest is ALWAYS a host pointer (4 words, one circuit per batch);
 * public_inputs_hashes (n_proofs x 4 words) lives where `mem` says, like the records.
 * Replaces: PlonkVerifierChip::get_challenges (chip/plonk/plonk_verifier_chip.rs:55-154). */
int sv_fri_challenges_batch(sv_ctx* ctx, const sv_fri_shape* shape, size_t n_proofs, uint64_t* records,
                            const uint64_t circuit_digest[4], const uint64_t* public_inputs_hashes,
                            uint32_t num_challenges, int mem);

/* get_challenges + verify_fri_proof in one call: the challenge fields of the records are ignored on
 * input and derived on the device before the query phase.  SV_MEM_DEVICE: the device records get their
 * challenge fields overwritten; SV_MEM_HOST: the host records are not modified.
 * Replaces: plonk_verifier_chip.rs:55-154 followed by chip/fri_chip.rs:329-362. */
int sv_fri_verify_batch_fs(sv_ctx* ctx, const sv_fri_shape* shape, size_t n_proofs, uint64_t* records,
                           const uint64_t circuit_digest[4], const uint64_t* public_inputs_hashes,
                           uint32_t num_challenges, uint32_t* accept_bitmap, uint32_t* first_fail, int mem);

/* Gather the per-rank accept bitmaps of a sharded batch: ncclAllGather(local -> full) on the ctx
 * stream.  nccl_comm is an ncclComm_t; libnccl.so.2 is resolved at run time (dlopen).  New -- the
 * reference is single-process. */
int sv_allgather_bitmap(sv_ctx* ctx, void* nccl_comm, const uint32_t* local_words, uint32_t* all_words,
                        size_t words_per_rank);

/* --- host side (CPU): transcript and synthetic proofs ----------------------------------------- */
/* Fiat-Shamir challenges of one proof, written into the record header (zeta, zeta_next, alpha,
 * betas, pow_response, indices).  Replaces: PlonkVerifierChip::get_challenges
 * (chip/plonk/plonk_verifier_chip.rs:55-154) + TranscriptChip + HasherChip::{update,squeeze}. */
int sv_fri_challenges(const sv_fri_shape* shape, uint64_t* record, const uint64_t circuit_digest[4],
                      const uint64_t public_inputs_hash[4], uint32_t num_challenges);

/* Synthetic plonky2-shaped proofs (valid by construction), `n_proofs` records written to
 * records_out.  `n_circuits` distinct committed oracle sets are built from `seed`; proof i uses
 * circuit i % n_circuits and its own public-input hash, so every proof has its own challenges,
 * query indices, openings and FRI commit phase.  Stand-in for plonky2's prover
 * (the plonky2_semaphore module, not on the hot path) -- test/bench data only. */
int sv_synth_proofs(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                    uint32_t num_challenges, uint64_t* records_out, int nthreads);

/* The same, with the public-inputs hash of every proof given by the caller (pi_hashes: n_proofs x 4 canonical words;
 * NULL = drawn from the seed like sv_synth_proofs), e.g. the hashes of public inputs that travel with the proofs in
 * the wire format. */
int sv_synth_proofs_pi(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                       uint32_t num_challenges, const uint64_t* pi_hashes, uint64_t* records_out, int nthreads);

/* The transcript inputs sv_synth_proofs used for the same (shape, seed, n_circuits, n_proofs):
 * circuit_digests_out: n_circuits x 4 words (proof i belongs to circuit i % n_circuits),
 * pi_hashes_out: n_proofs x 4 words.  Either pointer may be NULL. */
int sv_synth_public_inputs(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                           uint64_t* circuit_digests_out, uint64_t* pi_hashes_out);

/* --- wire format: plonky2 proof bytes -> flat records ---------------------------- */
/* The CommonCircuitData / CircuitConfig fields plonky2's `ProofWithPublicInputs::from_bytes(bytes, common_data)`
 * reads the vector lengths from (reference mirror: CommonData, types/common_data.rs:84-123; CircuitConfig :23-40). */
typedef struct sv_plonk_common {
    uint32_t num_constants;          /* CommonData.num_constants */
    uint32_t num_routed_wires;       /* CircuitConfig.num_routed_wires (= number of sigma polynomials) */
    uint32_t num_wires;              /* CircuitConfig.num_wires */
    uint32_t num_challenges;         /* CircuitConfig.num_challenges */
    uint32_t num_partial_products;   /* CommonData.num_partial_products */
    uint32_t quotient_degree_factor; /* CommonData.quotient_degree_factor */
    uint32_t num_public_inputs;      /* CommonData.num_public_inputs */
} sv_plonk_common;

/* first_fail code of a proof whose bytes do not parse (a Merkle-proof length byte that disagrees with the
 * shape); plonky2 would fail in from_bytes / verify_merkle_proof before any arithmetic */
#define SV_FAIL_MALFORMED 8
/* first_fail code of a proof whose openings do not satisfy the vanishing-polynomial identity (sv_verify_proofs_full) */
#define SV_FAIL_PLONK 9

/* FriParams/FriConfig + CommonData -> sv_fri_shape: the oracle widths and blinding flags of
 * CommonData::fri_oracles (types/common_data.rs:195-221; blinding = PlonkOracle consts :101-123), batch 1 =
 * fri_zs_polys (:176-178), final_poly_len = 2^(degree_bits - sum of arity bits).  reduction_arity_bits: num_steps entries
 * (NULL = arity 2 throughout, the reference's ConstantArityBits(1, 5) strategy). */
int sv_fri_shape_from_common(const sv_plonk_common* common, uint32_t degree_bits, uint32_t rate_bits, uint32_t cap_height,
                             uint32_t num_query_rounds, uint32_t proof_of_work_bits, uint32_t num_steps,
                             const uint32_t* reduction_arity_bits, uint32_t hiding, uint32_t hash_kind, sv_fri_shape* out);

/* Length in bytes of one serialised ProofWithPublicInputs of this shape (every proof of one circuit has the same
 * length); 0 if shape and common disagree.  Layout (plonky2 @ the revision Cargo.lock pins, util/serialization.rs,
 * write_proof_with_public_inputs; every field element a little-endian u64, a hash 4 of them, an extension element 2,
 * no length prefixes except one u8 per Merkle proof):
 *   wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap              3 x 2^cap_height x 32 B
 *   openings: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys   (x 16 B)
 *   commit_phase_merkle_caps                                                   num_steps x 2^cap_height x 32 B
 *   per query round: 4 x (leaf evals x 8 B, u8 n, n x 32 B siblings); per step (2^arity_bits x 16 B evals, u8 n, n x 32 B)
 *   final_poly coefficients x 16 B, pow_witness 8 B, public inputs x 8 B
 * Mirrors the field order of ProofValues / OpeningSetValues / FriProofValues (types/proof.rs:34-43,143-160,380-387). */
size_t sv_wire_proof_bytes(const sv_fri_shape* shape, const sv_plonk_common* common);

/* Host: record (+ public inputs) -> wire bytes; the inverse of sv_wire_unpack_batch, for fixtures, tests and the bench.
 * The challenge fields of the record are not part of the wire format.  bytes_out: sv_wire_proof_bytes bytes. */
int sv_wire_pack(const sv_fri_shape* shape, const sv_plonk_common* common, const uint64_t* record,
                 const uint64_t* public_inputs, uint8_t* bytes_out);

/* Host (CPU threads): n proofs at blob + i * stride_bytes -> records (challenge fields and padding zeroed, init_caps[0]
 * = constants_sigmas_cap of the verifier key, 2^cap_height x 4 words), the Poseidon-Goldilocks hash of each proof's
 * public inputs (PlonkVerifierChip::get_public_inputs_hash, plonk_verifier_chip.rs:41-53; pi_hashes_out n x 4 words,
 * may be NULL), the public inputs themselves (public_inputs_out n x num_public_inputs words, may be NULL) and a
 * malformed flag per proof (malformed_out n bytes, may be NULL; a malformed proof still yields a record).
 * Replaces: ProofWithPublicInputs::from_bytes + ProofValues::from (types/proof.rs:389-403) +
 * VerificationKeyValues::from (types/verification_key.rs:14-24) for the fields the FRI path reads. */
int sv_wire_unpack_batch(const sv_fri_shape* shape, const sv_plonk_common* common, const uint64_t* constants_sigmas_cap,
                         const uint8_t* blob, size_t stride_bytes, size_t n_proofs, uint64_t* records_out,
                         uint64_t* pi_hashes_out, uint64_t* public_inputs_out, uint8_t* malformed_out, int nthreads);

/* GPU: the same unpacking as one gather kernel (HBM-bound byte shuffling, one block per header / query round of a proof) plus the
 * public-input hashes, one thread per proof.  mem says where blob / records_out / pi_hashes_out / malformed_out
 * (n x u32) live; constants_sigmas_cap is always a host pointer.  SV_MEM_DEVICE: blob must be 8-byte aligned and
 * readable up to the next multiple of 8 bytes past its end. */
int sv_wire_unpack_batch_gpu(sv_ctx* ctx, const sv_fri_shape* shape, const sv_plonk_common* common,
                             const uint64_t* constants_sigmas_cap, const uint8_t* blob, size_t stride_bytes, size_t n_proofs,
                             uint64_t* records_out, uint64_t* pi_hashes_out, uint32_t* malformed_out, int mem);

/* The whole verifier-side path from serialised proofs in HOST memory: chunks of wire bytes move H2D back to back;
 * per chunk the device unpacks them into records, hashes the public inputs, derives the challenges (device
 * transcript) and runs the FRI query phase.  A malformed proof is rejected (bit 0, first_fail SV_FAIL_MALFORMED).
 * Replaces, per proof: from_bytes + ProofValues::from + get_public_inputs_hash + get_challenges + verify_fri_proof
 * (verifier_api.rs:34-56 down to chip/fri_chip.rs:329-362). */
int sv_verify_proofs_wire(sv_ctx* ctx, const sv_fri_shape* shape, const sv_plonk_common* common,
                          const uint64_t* constants_sigmas_cap, const uint64_t circuit_digest[4], const uint8_t* blob,
                          size_t stride_bytes, size_t n_proofs, uint32_t* accept_bitmap, uint32_t* first_fail);

/* hash_n_to_hash_no_pad over Poseidon-Goldilocks of n words (host): the public-inputs hash of one proof. */
int sv_public_inputs_hash(const uint64_t* public_inputs, size_t n, uint64_t out[4]);

/* --- plonk-level checks: the vanishing-polynomial identity at zeta ----------------- */
#define SV_GATE_NOOP 0         /* NoopGate, chip/plonk/gates/noop.rs */
#define SV_GATE_CONSTANT 1     /* ConstantGate { num_consts = param }, gates/constant.rs */
#define SV_GATE_PUBLIC_INPUT 2 /* PublicInputGate, gates/public_input.rs */
#define SV_GATE_ARITHMETIC 3   /* ArithmeticGate { num_ops = param }, gates/arithmetic.rs */
#define SV_GATE_ARITHMETIC_EXT 4 /* ArithmeticExtensionGate { num_ops = param }, gates/arithmetic_extension.rs */
#define SV_GATE_MUL_EXT 5      /* MulExtensionGate { num_ops = param }, gates/multiplication_extension.rs */
#define SV_GATE_BASE_SUM 6     /* BaseSumGate { num_limbs = param }, base 2, gates/base_sum.rs */
#define SV_GATE_REDUCING 7     /* ReducingGate { num_coeffs = param }, gates/reducing.rs */
#define SV_GATE_REDUCING_EXT 8 /* ReducingExtensionGate { num_coeffs = param }, gates/reducing_extension.rs */
#define SV_GATE_RANDOM_ACCESS 9 /* RandomAccessGate { bits = param, num_copies = param2, num_extra_constants = param3 }, gates/random_access.rs */
#define SV_GATE_POSEIDON_MDS 10 /* PoseidonMdsGate, gates/poseidon_mds.rs */
#define SV_GATE_POSEIDON 11    /* PoseidonGate (width 12), gates/poseidon.rs:324-700 */
/* = every gate the reference knows (gates/mod.rs:138-196); any other kind is refused by sv_plonk_circuit_check */
#define SV_MAX_GATES 32
#define SV_MAX_SELECTORS 8
#define SV_MAX_ROUTED_WIRES 128
#define SV_MAX_PLONK_CHALLENGES 4
#define SV_MAX_GATE_CONSTRAINTS 128

typedef struct sv_plonk_gate {
    uint32_t kind;           /* SV_GATE_* */
    uint32_t param;          /* num_consts / num_ops / num_limbs / num_coeffs / bits */
    uint32_t param2, param3; /* RandomAccessGate: num_copies, num_extra_constants; else 0 */
    uint32_t selector_index; /* SelectorsInfo.selector_indices[gate] (types/common_data.rs:56-66) */
} sv_plonk_gate;

/* What eval_vanishing_poly reads from CommonData (types/common_data.rs:68-96): gate list in CommonData.gates order
 * (the position of a gate is its selector value), selector groups, k_is, constraint count. */
typedef struct sv_plonk_circuit {
    sv_plonk_common common;
    uint32_t degree_bits;          /* FriParams.degree_bits: n = 2^degree_bits rows */
    uint32_t num_gate_constraints; /* CommonData.num_gate_constraints */
    uint32_t num_selectors;        /* SelectorsInfo.num_selectors() = groups.len(); the first num_selectors constants */
    uint32_t group_lo[SV_MAX_SELECTORS], group_hi[SV_MAX_SELECTORS]; /* SelectorsInfo.groups[s] = lo..hi */
    uint32_t num_gates;
    sv_plonk_gate gates[SV_MAX_GATES];
    uint64_t k_is[SV_MAX_ROUTED_WIRES]; /* CommonData.k_is, one coset shift per routed wire */
} sv_plonk_circuit;

/* plonky2 gate id string (`gate.0.id()`, the key CustomGateRef::from matches on, chip/plonk/gates/mod.rs:138-196) ->
 * kind and parameters; selector_index is left 0.  Accepts the ids of that table with ANY numeric parameters (the
 * reference hard-codes the ones of its recursion circuits), base-2 BaseSumGate only; < 0 for an unknown id
 * (the reference: unimplemented!()). */
int sv_plonk_gate_from_id(const char* gate_id, sv_plonk_gate* out);

/* What CommonData::from reads out of plonky2's CommonCircuitData (types/common_data.rs:224-270), as plain arrays: the Rust
 * side passes `gate.0.id()` of every gate (the string CustomGateRef::from matches, chip/plonk/gates/mod.rs:138-196),
 * SelectorsInfo (selector_indices per gate, groups as [start, end) ranges) and k_is. */
typedef struct sv_common_circuit_data {
    sv_plonk_common common;           /* CircuitConfig.num_wires / num_routed_wires / num_challenges, CommonCircuitData.num_constants /
                                         num_partial_products / quotient_degree_factor / num_public_inputs */
    uint32_t rate_bits, cap_height, proof_of_work_bits, num_query_rounds; /* config.fri_config */
    uint32_t hiding, degree_bits;     /* fri_params */
    uint32_t num_reduction_steps;
    const uint32_t* reduction_arity_bits; /* fri_params.reduction_arity_bits[num_reduction_steps] */
    uint32_t num_gate_constraints;
    uint32_t num_gates;
    const char* const* gate_ids;      /* gates[i].0.id() */
    const uint32_t* selector_indices; /* selectors_info.selector_indices[num_gates] */
    uint32_t num_selector_groups;
    const uint32_t* group_starts;     /* selectors_info.groups[s].start */
    const uint32_t* group_ends;       /* selectors_info.groups[s].end */
    uint32_t num_k_is;
    const uint64_t* k_is;             /* k_is[num_k_is], num_k_is = num_routed_wires */
    uint32_t hash_kind;               /* SV_HASH_*: the Hasher of the proof's GenericConfig */
} sv_common_circuit_data;

/* CommonCircuitData -> (sv_fri_shape, sv_plonk_circuit): gate ids -> kinds and parameters (an id outside the reference's table is
 * refused, like its unimplemented!()), selector groups, k_is, oracle widths; the result passed sv_plonk_circuit_check.
 * Replaces: CommonData::from (types/common_data.rs:224-270) + CustomGateRef::from (chip/plonk/gates/mod.rs:138-196). */
int sv_circuit_from_common_data(const sv_common_circuit_data* cd, sv_fri_shape* shape_out, sv_plonk_circuit* circuit_out);

/* 0 if the circuit description is consistent and uses only gates this library evaluates, else < 0. */
int sv_plonk_circuit_check(const sv_plonk_circuit* circuit);

/* Host: the plonk challenges of one proof -- out[0..nc) = plonk_betas, [nc..2nc) = plonk_gammas, [2nc..3nc) =
 * plonk_alphas -- from the same transcript as sv_fri_challenges (plonk_verifier_chip.rs:65-103), which squeezes and
 * drops them on its way to zeta. */
int sv_plonk_challenges(const sv_fri_shape* shape, const uint64_t* record, const uint64_t circuit_digest[4],
                        const uint64_t public_inputs_hash[4], uint32_t num_challenges, uint64_t* out);

/* For each of n_proofs records (openings at off_open0 / off_open1, zeta at off_zeta): does
 * vanishing_i(zeta) == Z_H(zeta) * sum_j quotient_{i,j}(zeta) * zeta^(n j) hold for every challenge i?  Writes
 * ceil(n/32) bitmap words (bit = identity holds; AND it with the FRI bitmap for the verdict of
 * verify_proof_with_challenges).  pi_hashes: n x 4 words; plonk_challenges: n x 3*num_challenges words
 * (sv_plonk_challenges).  A non-canonical input word fails the proof.
 * Replaces: PlonkVerifierChip::verify_proof_with_challenges up to the FRI call (plonk_verifier_chip.rs:156-210) with
 * eval_vanishing_poly (vanishing_poly.rs:18-218) and the gate constraints of every gate under chip/plonk/gates/.
 * sv_plonk_check_batch: one GPU thread per proof (mem = where records / pi_hashes / plonk_challenges / accept_bitmap
 * live); sv_plonk_check_host: the same function on CPU threads. */
int sv_plonk_check_batch(sv_ctx* ctx, const sv_fri_shape* shape, const sv_plonk_circuit* circuit, size_t n_proofs,
                         const uint64_t* records, const uint64_t* pi_hashes, const uint64_t* plonk_challenges,
                         uint32_t* accept_bitmap, int mem);
int sv_plonk_check_host(const sv_fri_shape* shape, const sv_plonk_circuit* circuit, size_t n_proofs, const uint64_t* records,
                        const uint64_t* pi_hashes, const uint64_t* plonk_challenges, uint32_t* accept_bitmap, int nthreads);

/* The complete verifier for a batch of serialised proofs of one circuit: sv_verify_proofs_wire plus the plonk-level
 * identity, all on the device -- unpack, public-inputs hash, plonk challenges, transcript, vanishing-polynomial check,
 * FRI query phase; bit i = proof i verifies.  first_fail: SV_FAIL_MALFORMED, then SV_FAIL_PLONK, then the FRI codes.
 * Replaces: verify_inside_snark_mock's whole verification (verifier_api.rs:34-56 -> Verifier::synthesize ->
 * PlonkVerifierChip::{get_public_inputs_hash, get_challenges, verify_proof_with_challenges}), natively and per batch. */
int sv_verify_proofs_full(sv_ctx* ctx, const sv_fri_shape* shape, const sv_plonk_circuit* circuit,
                          const uint64_t* constants_sigmas_cap, const uint64_t circuit_digest[4], const uint8_t* blob,
                          size_t stride_bytes, size_t n_proofs, uint32_t* accept_bitmap, uint32_t* first_fail);

/* --- commit-phase library: NTT / low-degree extension over Goldilocks ------------------ */
/* n_polys polynomials of 2^log_n words each, contiguous, transformed in place.
 * inverse = 0: coefficients (natural order) -> evaluations, the value at omega^bitrev(i) in position i (omega =
 * 7^((p-1)/2^log_n)) -- the order in which plonky2 puts evaluations into Merkle leaves and the FRI verifier reads them
 * back (chip/fri_chip.rs:152-166, 262-264).  inverse = 1: the reverse, 1/n included.
 * Replaces: plonky2_field's fft / ifft as used by the prover side of the reference (the plonky2_semaphore module, not on the
 * verifier's hot path); shared-memory tiled radix-8 rounds: 1 HBM pass up to 2^13 points, 2 up to 2^22. */
int sv_ntt_batch(sv_ctx* ctx, uint32_t log_n, size_t n_polys, uint64_t* data, int inverse, int mem);
/* out[p][i] = f_p(shift * omega_N^bitrev(i)), N = 2^(log_n + rate_bits): the low-degree extension of n_polys coefficient
 * vectors onto the coset shift * <omega_N>, in Merkle-leaf order (plonky2: PolynomialBatch::from_coeffs / lde_values with
 * shift = 7).  coeffs: n_polys x 2^log_n, out: n_polys x N. */
int sv_lde_batch(sv_ctx* ctx, uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint64_t shift,
                 uint64_t* out, int mem);
/* The prover's commitment to n_polys polynomials in one call: LDE onto 7 * <omega_N> (sv_lde_batch), leaves = one row of
 * n_polys values per evaluation point (a transposing copy), Merkle tree down to the cap (sv_merkle_tree_build).
 * leaves_out: N x n_polys words (NULL: not wanted -- the leaf digests are hashed straight from the polynomial-major values);
 * layers_out: 4 * (2N - 2^cap_height) words as in sv_merkle_tree_build.
 * Replaces: plonky2 PolynomialBatch::from_coeffs (LDE + MerkleTree::new), the commit step of the reference's prover side. */
int sv_commit_batch(sv_ctx* ctx, uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint32_t cap_height,
                    int hash_kind, uint64_t* leaves_out, uint64_t* layers_out, int mem);
/* the same two transforms on CPU threads (the function the kernels run, one "thread" per tile) */
int sv_ntt_host(uint32_t log_n, size_t n_polys, uint64_t* data, int inverse, int nthreads);
int sv_lde_host(uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint64_t shift, uint64_t* out,
                int nthreads);

/* library / build info */
const char* sv_version(void);

#ifdef __cplusplus
}
#endif
#endif
