/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h for the parity status).
 *
 * Statement-by-statement CPU restatement of the reference's FRI query-phase verifier.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference/src/plonky2_verifier/).  Written for clarity, not speed: `%` reductions on
 * unsigned __int128, canonical values everywhere.
 */
#include "oracle.h"
#include "orc_field.h"
#include "poseidon_g_constants.h"
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------
 * Poseidon over Goldilocks: width 12, S-box x^7, 4 + 22 + 4 rounds.
 * ---------------------------------------------------------------------------------------- */
/* All table entries are canonical (< p); orc_tables_canonical() lets the tests assert it. */
int orc_tables_canonical(void) {
    int ok = 1;
    for (int i = 0; i < 360; i++) ok &= ORC_ALL_ROUND_CONSTANTS[i] < ORC_P;
    for (int i = 0; i < 12; i++) ok &= ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i] < ORC_P;
    for (int i = 0; i < 22; i++) ok &= ORC_FAST_PARTIAL_ROUND_CONSTANTS[i] < ORC_P;
    for (int i = 0; i < 242; i++) ok &= ORC_FAST_PARTIAL_ROUND_VS[i] < ORC_P && ORC_FAST_PARTIAL_ROUND_W_HATS[i] < ORC_P;
    for (int i = 0; i < 121; i++) ok &= ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[i] < ORC_P;
    return ok;
}

static uint64_t sbox7(uint64_t x) { /* gates/poseidon.rs:428-436: exp(element, 7) */
    uint64_t x2 = orc_mul(x, x), x4 = orc_mul(x2, x2), x3 = orc_mul(x, x2);
    return orc_mul(x3, x4);
}

/* gates/poseidon.rs:450-502: row r of the MDS = sum_i CIRC[i]*state[(i+r)%12] + DIAG[r]*state[r] */
static void mds_layer(uint64_t st[12]) {
    uint64_t out[12];
    for (int r = 0; r < 12; r++) {
        orc_u128 acc = 0; /* 12 * 41 * 2^64 < 2^128: exact */
        for (int i = 0; i < 12; i++) acc += (orc_u128)ORC_MDS_MATRIX_CIRC[i] * st[i + r < 12 ? i + r : i + r - 12];
        acc += (orc_u128)ORC_MDS_MATRIX_DIAG[r] * st[r];
        out[r] = orc_red128(acc);
    }
    memcpy(st, out, sizeof out);
}

static void full_round(uint64_t st[12], int round_ctr) {
    for (int i = 0; i < 12; i++) st[i] = orc_add(st[i], ORC_ALL_ROUND_CONSTANTS[i + 12 * round_ctr]); /* :383-406 */
    for (int i = 0; i < 12; i++) st[i] = sbox7(st[i]);                                                     /* :438-448 */
    mds_layer(st);
}

/* Naive definition: 30 x (add constants, S-box on all lanes (full) or lane 0 (partial), MDS).
 * Constants: ALL_ROUND_CONSTANTS (== chip/poseidon_spec/constants.rs:7-443). */
void orc_poseidon_naive(uint64_t st[12]) {
    for (int i = 0; i < 12; i++) st[i] %= ORC_P;
    int rc = 0;
    for (int r = 0; r < 4; r++) full_round(st, rc++);
    for (int r = 0; r < 22; r++) {
        for (int i = 0; i < 12; i++) st[i] = orc_add(st[i], ORC_ALL_ROUND_CONSTANTS[i + 12 * rc]);
        st[0] = sbox7(st[0]);
        mds_layer(st);
        rc++;
    }
    for (int r = 0; r < 4; r++) full_round(st, rc++);
}

/* Fast form, the one the reference restates in-circuit: gates/poseidon.rs:634-686. */
void orc_poseidon(uint64_t st[12]) {
    for (int i = 0; i < 12; i++) st[i] %= ORC_P;
    int rc = 0;
    for (int r = 0; r < 4; r++) full_round(st, rc++);                               /* :637-650 */
    for (int i = 0; i < 12; i++) st[i] = orc_add(st[i], ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]); /* :652, :408-426 */
    {   /* mds_partial_layer_init :504-537 */
        uint64_t res[12] = {0};
        res[0] = st[0];
        for (int c = 1; c < 12; c++) {                  /* column c: sum_r M[r-1][c-1] * st[r], reduced once */
            orc_u128 acc = 0;
            uint64_t wraps = 0;
            for (int r = 1; r < 12; r++) {
                orc_u128 t = (orc_u128)ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)] * st[r], s2 = acc + t;
                wraps += s2 < t;
                acc = s2;
            }
            res[c] = orc_sub(orc_red128(acc), wraps << 32);
        }
        memcpy(st, res, sizeof res);
    }
    for (int r = 0; r < 22; r++) {                                                   /* :654-672 */
        st[0] = sbox7(st[0]);
        if (r != 21) st[0] = orc_add(st[0], ORC_FAST_PARTIAL_ROUND_CONSTANTS[r]);
        /* mds_partial_layer_fast :539-589.  The 12-term dot product is summed exactly (128-bit accumulator plus a
         * count of its wrap-arounds) and reduced once: 2^128 = (2^64)^2 = (2^32 - 1)^2 = -2^32 (mod p). */
        orc_u128 acc = (orc_u128)st[0] * (ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0]);
        uint64_t wraps = 0;
        for (int i = 1; i < 12; i++) {
            orc_u128 t = (orc_u128)ORC_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1] * st[i], s2 = acc + t;
            wraps += s2 < t;
            acc = s2;
        }
        uint64_t d = orc_sub(orc_red128(acc), wraps << 32);
        uint64_t res[12];
        res[0] = d;
        for (int i = 1; i < 12; i++) res[i] = orc_mul_add(ORC_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1], st[0], st[i]);
        memcpy(st, res, sizeof res);
    }
    rc += 22;                                                                        /* :673 */
    for (int r = 0; r < 4; r++) full_round(st, rc++);                               /* :675-686 */
}


/* ------------------------------------------------------------------------------------------
 * Hash family B: Poseidon over the BN254 scalar field wrapped around 12 Goldilocks limbs.
 *   parameters ........ bn245_poseidon/constants.rs:5-384 (340 round constants, 5x5 MDS), :402-404
 *                       (T = 5, R_F = 8, R_P = 60)
 *   permutation ....... bn245_poseidon/native.rs:16-60 (constant layer, x^5 on all lanes / lane 0, MDS
 *                       new[i] = sum_j state[j] * M[i][j])
 *   encode / decode ... native.rs:62-77 (3 limbs -> sum x_i p^i; 4 base-p digits, keep 3),
 *                       native_chip/utils.rs:25-36 (goldilocks_decompose)
 *   wrapper ........... bn245_poseidon/plonky2_config.rs:38-51 (4 chunks of 3, pad to 5 with 0,
 *                       decode the first 4 words), in-circuit twin native_chip/all_chip.rs:52-89
 * Fr arithmetic lives in halo2curves bn256::Fr 0.3.2 (Cargo.lock:1063-1065, not on disk); this restates
 * its published representation: 4 x 64-bit limbs, Montgomery multiplication (CIOS) with R = 2^256.
 * Pinned by the SURVEY 8c KATs and by tests/golden/poseidon_b.json (pure-Python big integers).
 * ---------------------------------------------------------------------------------------- */
#include "poseidon_b_constants.h"
typedef struct { uint64_t l[4]; } orc_fr;
static const uint64_t FR_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_NINV = 0xc2e1f593efffffffULL;  /* -r^-1 mod 2^64 */
static const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};

static int fr_geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > FR_MOD[i]) return 1; if (a[i] < FR_MOD[i]) return 0; }
    return 1;
}
static void fr_sub_mod(uint64_t a[4]) {
    orc_u128 br = 0;
    for (int i = 0; i < 4; i++) { orc_u128 d = (orc_u128)a[i] - FR_MOD[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
static orc_fr fr_add(orc_fr a, orc_fr b) {
    orc_fr r; orc_u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (orc_u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || fr_geq_mod(r.l)) fr_sub_mod(r.l);   /* a + b < 2r < 2^255: c is always 0 */
    return r;
}
/* Montgomery product a*b*2^-256 mod r (CIOS, Koc et al.) */
static orc_fr fr_mmul(orc_fr a, orc_fr b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        orc_u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (orc_u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_NINV;
        c = (orc_u128)m * FR_MOD[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (orc_u128)m * FR_MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    orc_fr r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fr_geq_mod(r.l)) fr_sub_mod(r.l);
    return r;
}
static orc_fr fr_to_mont(orc_fr a) { orc_fr r2 = {{FR_R2[0], FR_R2[1], FR_R2[2], FR_R2[3]}}; return fr_mmul(a, r2); }
static orc_fr fr_from_mont(orc_fr a) { orc_fr one = {{1, 0, 0, 0}}; return fr_mmul(a, one); }

static orc_fr B_RC[340], B_MDS[25];
static pthread_once_t b_once = PTHREAD_ONCE_INIT;
static void b_init(void) {
    for (int i = 0; i < 340; i++) { orc_fr v; memcpy(v.l, ORC_B_ROUND_CONSTANTS + 4 * i, 32); B_RC[i] = fr_to_mont(v); }
    for (int i = 0; i < 25; i++) { orc_fr v; memcpy(v.l, ORC_B_MDS + 4 * i, 32); B_MDS[i] = fr_to_mont(v); }
}
static orc_fr fr_pow5(orc_fr x) { orc_fr x2 = fr_mmul(x, x), x4 = fr_mmul(x2, x2); return fr_mmul(x4, x); }
/* native.rs:43-60, state in Montgomery form */
static void b_permute_mont(orc_fr st[5]) {
    pthread_once(&b_once, b_init);
    int counter = 0;
    for (int round = 0; round < 68; round++) {
        for (int i = 0; i < 5; i++) st[i] = fr_add(st[i], B_RC[counter++]);          /* constant_layer :16-21 */
        if (round < 4 || round >= 64) { for (int i = 0; i < 5; i++) st[i] = fr_pow5(st[i]); } /* sbox_layer :23-27 */
        else st[0] = fr_pow5(st[0]);                                                  /* partial_sbox_layer :29-31 */
        orc_fr nw[5];                                                                 /* mds_layer :33-41 */
        for (int i = 0; i < 5; i++) {
            orc_fr acc = {{0, 0, 0, 0}};
            for (int j = 0; j < 5; j++) acc = fr_add(acc, fr_mmul(st[j], B_MDS[5 * i + j]));
            nw[i] = acc;
        }
        memcpy(st, nw, sizeof nw);
    }
}
/* raw Fr permutation on canonical integers (KAT entry point): state[5][4] little-endian limbs */
void orc_poseidon_b_fr(uint64_t state[20]) {
    orc_fr st[5];
    for (int i = 0; i < 5; i++) { memcpy(st[i].l, state + 4 * i, 32); st[i] = fr_to_mont(st[i]); }
    b_permute_mont(st);
    for (int i = 0; i < 5; i++) { st[i] = fr_from_mont(st[i]); memcpy(state + 4 * i, st[i].l, 32); }
}
/* a[4] = a * m + add  (a < 2^192 stays < 2^256) */
static void limbs_mul_add(uint64_t a[4], uint64_t m, uint64_t add) {
    orc_u128 c = add;
    for (int i = 0; i < 4; i++) { c += (orc_u128)a[i] * m; a[i] = (uint64_t)c; c >>= 64; }
}
/* a[4] /= d, returns a % d (schoolbook long division, 128/64 steps) */
static uint64_t limbs_divrem(uint64_t a[4], uint64_t d) {
    orc_u128 rem = 0;
    for (int i = 3; i >= 0; i--) { orc_u128 cur = (rem << 64) | a[i]; a[i] = (uint64_t)(cur / d); rem = cur % d; }
    return (uint64_t)rem;
}
/* Bn254PoseidonPermutation::permute (plonky2_config.rs:38-51) */
void orc_poseidon_b(uint64_t st[12]) {
    orc_fr s5[5];
    for (int k = 0; k < 4; k++) {                       /* encode_fe: x0 + x1 p + x2 p^2 (native.rs:62-67) */
        uint64_t v[4] = {st[3 * k + 2] % ORC_P, 0, 0, 0};
        limbs_mul_add(v, ORC_P, st[3 * k + 1] % ORC_P);
        limbs_mul_add(v, ORC_P, st[3 * k] % ORC_P);
        memcpy(s5[k].l, v, 32);
        s5[k] = fr_to_mont(s5[k]);
    }
    memset(&s5[4], 0, sizeof s5[4]);                    /* resize(T, 0) */
    b_permute_mont(s5);
    for (int k = 0; k < 4; k++) {                       /* decode_fe: 4 base-p digits, keep 3 (native.rs:69-77) */
        orc_fr c = fr_from_mont(s5[k]);
        for (int i = 0; i < 3; i++) st[3 * k + i] = limbs_divrem(c.l, ORC_P);
    }
}

/* the permutation of the hash family in use (orc_shape.hash_kind: 0 = Poseidon-Goldilocks, 1 = B) */
static _Thread_local int cur_kind = 0;
static void permute_cur(uint64_t st[12]) { if (cur_kind == 1) orc_poseidon_b(st); else orc_poseidon(st); }
void orc_set_hash_kind(int kind) { cur_kind = kind; }

/* thin exports of the field restatement, for the KAT tests */
uint64_t orc_f_mul(uint64_t a, uint64_t b) { return orc_mul(a % ORC_P, b % ORC_P); }
uint64_t orc_f_pow(uint64_t a, uint64_t e) { return orc_pow(a % ORC_P, e); }
uint64_t orc_f_inv(uint64_t a) { return orc_inv(a % ORC_P); }
uint64_t orc_f_red128(uint64_t lo, uint64_t hi, int slow) {
    orc_u128 x = ((orc_u128)hi << 64) | lo;
    return slow ? orc_red128_slow(x) : orc_red128(x);
}
void orc_f2_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]) {
    orc_fp2 r = orc2_mul(orc2(a[0], a[1]), orc2(b[0], b[1])); out[0] = r.c[0]; out[1] = r.c[1];
}
void orc_f2_inv(const uint64_t a[2], uint64_t out[2]) {
    orc_fp2 r = orc2_inv(orc2(a[0], a[1])); out[0] = r.c[0]; out[1] = r.c[1];
}

void orc_poseidon_batch(uint64_t *states, size_t n) {
    for (size_t i = 0; i < n; i++) orc_poseidon(states + 12 * i);
}

/* HasherChip::hash (hasher_chip.rs:122-148): fresh zero state (:36-39); per 8-chunk overwrite the
 * first len lanes then permute; output = state[0..4]. */
void orc_hash_no_pad(const uint64_t *in, size_t n, uint64_t out[4]) {
    uint64_t st[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < len; i++) st[i] = in[off + i];
        permute_cur(st);
    }
    memcpy(out, st, 4 * sizeof(uint64_t));
}

/* HasherChip::permute on a fresh hasher (hasher_chip.rs:150-171; fresh per level at
 * merkle_proof_chip.rs:59): state = [l, r, 0,0,0,0], permute, take 4. */
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[12] = {0};
    memcpy(st, l, 32);
    memcpy(st + 4, r, 32);
    permute_cur(st);
    memcpy(out, st, 32);
}

/* MerkleProofChip::verify_merkle_proof_to_cap_with_cap_index (merkle_proof_chip.rs:39-87). */
int orc_merkle_verify(const uint64_t *leaf, size_t leaf_len, uint64_t index, const uint64_t *siblings,
                      size_t depth, const uint64_t *cap, size_t cap_index) {
    uint64_t state[4] = {0, 0, 0, 0};
    if (leaf_len <= 4) memcpy(state, leaf, leaf_len * 8);              /* :52-53 (hash_or_noop) */
    else orc_hash_no_pad(leaf, leaf_len, state);                        /* :55 */
    for (size_t lvl = 0; lvl < depth; lvl++) {                          /* :58 zip(bits, siblings) */
        const uint64_t *sib = siblings + 4 * lvl;
        int bit = (int)((index >> lvl) & 1);
        uint64_t nxt[4];
        if (bit) orc_two_to_one(sib, state, nxt);                       /* :61-69 select(sibling,state,bit) */
        else orc_two_to_one(state, sib, nxt);
        memcpy(state, nxt, 32);
    }
    for (int i = 0; i < 4; i++)                                         /* :73-84 */
        if (cap[4 * cap_index + i] != state[i]) return 0;
    return 1;
}

void orc_merkle_verify_batch(const uint64_t *records, size_t n, size_t leaf_len, size_t depth,
                             const uint64_t *indices, const uint64_t *caps, size_t cap_height, uint8_t *ok) {
    size_t stride = leaf_len + 4 * depth;
    for (size_t i = 0; i < n; i++) {
        const uint64_t *rec = records + i * stride;
        int canon = 1;
        for (size_t w = 0; w < stride; w++) canon &= rec[w] < ORC_P;
        uint64_t idx = indices[i];
        size_t cap_index = (size_t)(idx >> depth) & (((size_t)1 << cap_height) - 1);
        ok[i] = (uint8_t)(canon && orc_merkle_verify(rec, leaf_len, idx, rec + leaf_len, depth, caps, cap_index));
    }
}

/* ------------------------------------------------------------------------------------------
 * Flat record layout (independent restatement of include/stark_verifier_b200.h's rules):
 * every segment starts on a 4-word (32 B) boundary.
 * ---------------------------------------------------------------------------------------- */
static uint32_t up4(uint32_t x) { return (x + 3u) & ~3u; }

int orc_make_layout(const orc_shape *s, orc_layout *L) {
    memset(L, 0, sizeof *L);
    if (s->num_steps > 32 || s->cap_height > 16) return -1;
    L->ncap = 1u << s->cap_height;
    L->lde_bits = s->degree_bits + s->rate_bits;
    {
        uint32_t total = 0;
        for (uint32_t i = 0; i < s->num_steps; i++) {
            if (s->reduction_arity_bits[i] < 1 || s->reduction_arity_bits[i] > 4) return -1;
            total += s->reduction_arity_bits[i];
        }
        if (L->lde_bits < s->cap_height + total) return -1;
    }
    L->n0 = s->oracle_num_polys[0] + s->oracle_num_polys[1] + s->oracle_num_polys[2] + s->oracle_num_polys[3];
    L->n1 = s->num_zs;
    uint32_t o = 0;
    L->off_init_caps = o;   o = up4(o + 4 * L->ncap * 4);
    L->off_step_caps = o;   o = up4(o + s->num_steps * L->ncap * 4);
    L->off_open0 = o;       o = up4(o + 2 * L->n0);
    L->off_open1 = o;       o = up4(o + 2 * L->n1);
    L->off_final_poly = o;  o = up4(o + 2 * s->final_poly_len);
    L->off_pow_witness = o; o = up4(o + 1);
    L->off_alpha = o;       o = up4(o + 2);
    L->off_betas = o;       o = up4(o + 2 * s->num_steps);
    L->off_pow_response = o;o = up4(o + 1);
    L->off_indices = o;     o = up4(o + s->num_query_rounds);
    L->off_zeta = o;        o = up4(o + 2);
    L->off_zeta_next = o;   o = up4(o + 2);
    L->header_words = o;
    uint32_t q = 0;
    L->init_depth = L->lde_bits - s->cap_height;
    for (int k = 0; k < 4; k++) {
        L->leaf_len[k] = s->oracle_num_polys[k] + ((s->hiding && s->oracle_blinding[k]) ? 4u : 0u);
        L->q_off_init_evals[k] = q; q = up4(q + L->leaf_len[k]);
        L->q_off_init_sibs[k] = q;  q = up4(q + 4 * L->init_depth);
    }
    uint32_t consumed = 0;
    for (uint32_t i = 0; i < s->num_steps; i++) {
        consumed += s->reduction_arity_bits[i];
        L->step_depth[i] = L->lde_bits - consumed - s->cap_height;   /* the tree over the cosets of layer i */
        L->q_off_step_evals[i] = q; q = up4(q + (2u << s->reduction_arity_bits[i]));   /* 2^arity_bits Fp2 evals */
        L->q_off_step_sibs[i] = q;  q = up4(q + 4 * L->step_depth[i]);
    }
    L->query_words = q;
    L->record_words = L->header_words + s->num_query_rounds * q;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FRI verifier
 * ---------------------------------------------------------------------------------------- */
/* reduce_extension (goldilocks_extension_chip.rs:331-342): fold from the LAST term,
 * acc = acc*base + term. */
static orc_fp2 reduce_ext(orc_fp2 base, const uint64_t *terms_fp2, size_t n) {
    orc_fp2 acc = orc2(0, 0);
    for (size_t i = n; i-- > 0;) acc = orc2_mul_add(acc, base, orc2(terms_fp2[2 * i], terms_fp2[2 * i + 1]));
    return acc;
}
/* exp (goldilocks_extension_chip.rs:264-282) */
static orc_fp2 ext_exp(orc_fp2 b, size_t power) {
    orc_fp2 r = orc2(1, 0);
    for (size_t i = 0; i < power; i++) r = orc2_mul(r, b);
    return r;
}
/* exp_from_bits (goldilocks_chip.rs:388-404): prod over set bits i of base^(2^i) */
static uint64_t exp_from_bits(uint64_t base, const int *bits, size_t n) {
    uint64_t x = 1;
    for (size_t i = 0; i < n; i++) {
        uint64_t b = orc_pow(base, (uint64_t)1 << i);
        if (bits[i]) x = orc_mul(x, b);
    }
    return x;
}

static int all_canonical(const uint64_t *w, size_t n) {
    for (size_t i = 0; i < n; i++) if (w[i] >= ORC_P) return 0;
    return 1;
}

/* oracle/fast/fast_avx512.c (bench-only) includes this file and verifies the Merkle proofs of a proof 8 at a time with
 * AVX-512; it then runs check_consistency with this flag set for everything else.  Never set by the oracle itself. */
static _Thread_local int orc_merkle_elsewhere = 0;
static int merkle_step(const uint64_t *leaf, size_t leaf_len, uint64_t index, const uint64_t *siblings, size_t depth,
                       const uint64_t *cap, size_t cap_index) {
    return orc_merkle_elsewhere ? 1 : orc_merkle_verify(leaf, leaf_len, index, siblings, depth, cap, cap_index);
}

/* check_consistency (fri_chip.rs:228-327) for one query round.  returns fail code. */
static int check_consistency(const orc_shape *s, const orc_layout *L, const uint64_t *rec, uint32_t round,
                             const orc_fp2 reduced_openings[2]) {
    const uint64_t *q = rec + L->header_words + (size_t)round * L->query_words;
    const uint32_t lde_bits = L->lde_bits;
    const orc_fp2 alpha = orc2(rec[L->off_alpha], rec[L->off_alpha + 1]);

    /* Range check of every witness of this round (assign_value, arithmetic_chip.rs:256-268) */
    for (int k = 0; k < 4; k++) {
        if (!all_canonical(q + L->q_off_init_evals[k], L->leaf_len[k])) return ORC_FAIL_NONCANONICAL;
        if (!all_canonical(q + L->q_off_init_sibs[k], 4 * L->init_depth)) return ORC_FAIL_NONCANONICAL;
    }
    for (uint32_t i = 0; i < s->num_steps; i++) {
        if (!all_canonical(q + L->q_off_step_evals[i], 2u << s->reduction_arity_bits[i])) return ORC_FAIL_NONCANONICAL;
        if (!all_canonical(q + L->q_off_step_sibs[i], 4 * L->step_depth[i])) return ORC_FAIL_NONCANONICAL;
    }

    /* :245-250  x_index_bits = to_bits(x_index, 64).take(lde_bits) */
    uint64_t x_index_fe = rec[L->off_indices + round];
    int bits[64];
    for (int i = 0; i < 64; i++) bits[i] = (int)((x_index_fe >> i) & 1);
    uint32_t nbits = lde_bits;
    /* :72-82, :252  cap_index = from_bits(x_index_bits[len - cap_height ..]) */
    size_t cap_index = 0;
    for (uint32_t i = 0; i < s->cap_height; i++) cap_index |= (size_t)bits[nbits - s->cap_height + i] << i;
    uint64_t x_index = x_index_fe & (((uint64_t)1 << lde_bits) - 1);

    /* :254-260, :85-110 verify_initial_merkle_proof */
    for (int k = 0; k < 4; k++) {
        const uint64_t *cap = rec + L->off_init_caps + (size_t)k * L->ncap * 4;
        if (!merkle_step(q + L->q_off_init_evals[k], L->leaf_len[k], x_index,
                               q + L->q_off_init_sibs[k], L->init_depth, cap, cap_index))
            return ORC_FAIL_INIT_MERKLE;
    }

    /* :262-264, :152-166  x = offset * omega^{rev(bits)};  offset = MULTIPLICATIVE_GROUP_GENERATOR = 7
     * (plonk_verifier_chip.rs:225-227), omega = 7^((p-1)/2^lde_bits) */
    uint64_t omega = orc_pow(7, (ORC_P - 1) >> lde_bits);
    int rev[64];
    for (uint32_t i = 0; i < nbits; i++) rev[i] = bits[nbits - 1 - i];
    uint64_t x = orc_mul(7, exp_from_bits(omega, rev, nbits));

    /* :266-273, :112-149 batch_initial_polynomials */
    orc_fp2 sum = orc2(0, 0);
    for (int b = 0; b < 2; b++) {
        /* evals of the batch, in FriInstanceInfo order (fri.rs:50-72, common_data.rs:192-221);
         * unsalted_eval drops the 4 salt limbs at the END of a blinded leaf (assigned.rs:57-71),
         * which never shifts a polynomial index. */
        uint64_t evals[2 * 1024];
        size_t n = 0;
        if (b == 0) {
            for (int k = 0; k < 4; k++)
                for (uint32_t j = 0; j < s->oracle_num_polys[k]; j++) { evals[2 * n] = q[L->q_off_init_evals[k] + j]; evals[2 * n + 1] = 0; n++; }
        } else {
            for (uint32_t j = 0; j < s->num_zs; j++) { evals[2 * n] = q[L->q_off_init_evals[2] + j]; evals[2 * n + 1] = 0; n++; }
        }
        orc_fp2 point = b == 0 ? orc2(rec[L->off_zeta], rec[L->off_zeta + 1]) : orc2(rec[L->off_zeta_next], rec[L->off_zeta_next + 1]);
        orc_fp2 reduced_evals = reduce_ext(alpha, evals, n);                 /* :139-140 */
        orc_fp2 numerator = orc2_sub(reduced_evals, reduced_openings[b]);     /* :141-142 */
        orc_fp2 denominator = orc2_sub(orc2(x, 0), point);                    /* :143 */
        sum = orc2_mul(ext_exp(alpha, n), sum);                               /* :144 shift */
        if (orc2_is_zero(denominator)) return ORC_FAIL_ZERO_DENOM;            /* div_extension :78-81,:98-99 */
        sum = orc2_add(orc2_mul(numerator, orc2_inv(denominator)), sum);      /* :145-146 */
    }
    orc_fp2 prev_eval = sum;

    /* :275-316 */
    const int *xb = bits; /* current x_index_bits window */
    uint32_t xb_len = nbits;
    for (uint32_t i = 0; i < s->num_steps; i++) {
        const uint32_t arity_bits = s->reduction_arity_bits[i], arity = 1u << arity_bits;
        const uint64_t *evals = q + L->q_off_step_evals[i];      /* arity Fp2 */
        const int *coset_index_bits = xb + arity_bits;           /* :279 */
        uint32_t coset_len = xb_len - arity_bits;
        uint32_t x_index_within_coset = 0;                       /* :280-282 from_bits(x_index_bits[..arity_bits]) */
        for (uint32_t j = 0; j < arity_bits; j++) x_index_within_coset |= (uint32_t)xb[j] << j;
        /* :285-292 evals[x_index_within_coset] == prev_eval, limb-wise */
        if (evals[2 * x_index_within_coset] != prev_eval.c[0] || evals[2 * x_index_within_coset + 1] != prev_eval.c[1])
            return ORC_FAIL_STEP_EVAL;
        /* next_eval :168-226.  The reference stops at arity 2 (:211 TODO); the general case is plonky2's
         * compute_evaluation (fri/verifier.rs of the pinned dependency), which next_eval restates line by line up to the
         * interpolation: g = primitive arity-th root; reverse_index_bits(evals); coset_start = x * g^-rev(within);
         * interpolate {(coset_start * g^j, evals_rev[j])} and evaluate at beta (barycentric form). */
        {
            uint64_t g = orc_pow(7, (ORC_P - 1) >> arity_bits);   /* :181-184 */
            uint64_t g_inv = orc_inv(g);
            orc_fp2 ev[16];
            for (uint32_t j = 0; j < arity; j++) {                /* reverse_index_bits_in_place :188-189 */
                uint32_t r = 0;
                for (uint32_t b = 0; b < arity_bits; b++) r |= ((j >> b) & 1u) << (arity_bits - 1 - b);
                ev[j] = orc2(evals[2 * r], evals[2 * r + 1]);
            }
            int rb[4];
            for (uint32_t b = 0; b < arity_bits; b++) rb[b] = xb[arity_bits - 1 - b];   /* bits reversed :191-199 */
            uint64_t start = exp_from_bits(g_inv, rb, arity_bits);
            uint64_t coset_start = orc_mul(start, x);             /* :200 */
            uint64_t px[16];                                      /* points :203-210 */
            uint64_t g_power = 1;
            for (uint32_t j = 0; j < arity; j++) { px[j] = orc_mul(coset_start, g_power); g_power = orc_mul(g_power, g); }
            orc_fp2 beta = orc2(rec[L->off_betas + 2 * i], rec[L->off_betas + 2 * i + 1]);
            if (arity == 2) {
                /* the reference's literal two-point formula :212-224 */
                orc_fp2 a0 = orc2(px[0], 0), a1 = ev[0], b0 = orc2(px[1], 0), b1 = ev[1];
                orc_fp2 numerator = orc2_mul(orc2_sub(beta, a0), orc2_sub(b1, a1));   /* :219-221 */
                orc_fp2 denominator = orc2_sub(b0, a0);                               /* :222 */
                if (orc2_is_zero(denominator)) return ORC_FAIL_ZERO_DENOM;
                prev_eval = orc2_add(orc2_mul(numerator, orc2_inv(denominator)), a1); /* :223-224 */
            } else {
                /* plonky2 interpolate(points, beta, barycentric_weights(points)):
                 *   w_j = 1 / prod_{m != j} (x_j - x_m);  if beta == x_j return y_j;
                 *   l(beta) = prod_j (beta - x_j);  result = l(beta) * sum_j w_j / (beta - x_j) * y_j */
                int hit = -1;
                for (uint32_t j = 0; j < arity; j++) if (beta.c[1] == 0 && beta.c[0] == px[j]) hit = (int)j;
                if (hit >= 0) prev_eval = ev[hit];
                else {
                    orc_fp2 l = orc2(1, 0), sum = orc2(0, 0);
                    for (uint32_t j = 0; j < arity; j++) l = orc2_mul(l, orc2_sub(beta, orc2(px[j], 0)));
                    for (uint32_t j = 0; j < arity; j++) {
                        uint64_t d = 1;
                        for (uint32_t m = 0; m < arity; m++) if (m != j) d = orc_mul(d, orc_sub(px[j], px[m]));
                        orc_fp2 w_over = orc2_mul(orc2(orc_inv(d), 0), orc2_inv(orc2_sub(beta, orc2(px[j], 0))));
                        sum = orc2_add(sum, orc2_mul(w_over, ev[j]));
                    }
                    prev_eval = orc2_mul(l, sum);
                }
            }
        }
        /* :302-311 step Merkle proof: leaf = flattened evals (2 * arity limbs), index = coset_index_bits,
         * SAME cap_index as the initial trees (:308) */
        uint64_t coset_index = 0;
        for (uint32_t j = 0; j < coset_len; j++) coset_index |= (uint64_t)coset_index_bits[j] << j;
        const uint64_t *cap = rec + L->off_step_caps + (size_t)i * L->ncap * 4;
        if (!merkle_step(evals, 2 * arity, coset_index, q + L->q_off_step_sibs[i], L->step_depth[i], cap, cap_index))
            return ORC_FAIL_STEP_MERKLE;
        for (uint32_t b = 0; b < arity_bits; b++) x = orc_mul(x, x);   /* :313 exp_power_of_2(x, arity_bits) */
        xb = coset_index_bits; xb_len = coset_len;                /* :315 */
    }
    /* :317-325 final_poly(x) == prev_eval; reduce_extension_field_terms_base
     * (goldilocks_extension_chip.rs:357-365) = Horner with base (x,0) */
    orc_fp2 final_eval = reduce_ext(orc2(x, 0), rec + L->off_final_poly, s->final_poly_len);
    if (!orc2_eq(prev_eval, final_eval)) return ORC_FAIL_FINAL;
    return ORC_OK;
}

int orc_fri_verify(const orc_shape *s, const uint64_t *rec, int *fail, int *fail_query) {
    orc_layout L;
    int code = ORC_OK, fq = -1;
    if (orc_make_layout(s, &L)) { if (fail) *fail = -1; return 0; }
    cur_kind = (int)s->hash_kind;
    /* range check of the per-proof witnesses */
    if (!all_canonical(rec, L.header_words)) { code = ORC_FAIL_NONCANONICAL; goto done; }
    /* fri_verify_proof_of_work (fri_chip.rs:364-376): top proof_of_work_bits of the 64-bit
     * decomposition of the canonical pow_response are zero */
    {
        uint64_t r = rec[L.off_pow_response];
        for (uint32_t i = 0; i < s->proof_of_work_bits; i++)
            if ((r >> (63 - i)) & 1) { code = ORC_FAIL_POW; goto done; }
    }
    {
        /* compute_reduced_openings (fri_chip.rs:58-70) */
        orc_fp2 alpha = orc2(rec[L.off_alpha], rec[L.off_alpha + 1]);
        orc_fp2 ro[2];
        ro[0] = reduce_ext(alpha, rec + L.off_open0, L.n0);
        ro[1] = reduce_ext(alpha, rec + L.off_open1, L.n1);
        for (uint32_t r = 0; r < s->num_query_rounds; r++) {          /* :348-361 */
            code = check_consistency(s, &L, rec, r, ro);
            if (code != ORC_OK) { fq = (int)r; break; }
        }
    }
done:
    if (fail) *fail = code;
    if (fail_query) *fail_query = fq;
    return code == ORC_OK;
}

typedef struct { const orc_shape *s; const uint64_t *records; size_t n, stride; uint8_t *ok; int tid, nt; } batch_arg;
static void *batch_worker(void *p) {
    batch_arg *a = (batch_arg *)p;
    /* proofs are dealt round-robin; every thread writes its own bytes, the caller packs the bitmap */
    for (size_t i = (size_t)a->tid; i < a->n; i += (size_t)a->nt)
        a->ok[i] = (uint8_t)orc_fri_verify(a->s, a->records + i * a->stride, NULL, NULL);
    return NULL;
}
void orc_fri_verify_batch(const orc_shape *s, const uint64_t *records, size_t n, uint32_t *bitmap, int nthreads) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    uint8_t *ok = (uint8_t *)calloc(n ? n : 1, 1);
    pthread_t th[256];
    batch_arg args[256];
    for (int t = 0; t < nthreads; t++) {
        args[t] = (batch_arg){s, records, n, L.record_words, ok, t, nthreads};
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    for (size_t w = 0; w < (n + 31) / 32; w++) {
        uint32_t word = 0;
        for (size_t i = w * 32; i < n && i < w * 32 + 32; i++) word |= (uint32_t)ok[i] << (i & 31);
        bitmap[w] = word;
    }
    free(ok);
}

/* ------------------------------------------------------------------------------------------
 * Duplex challenger: HasherChip::{update,absorb_buffered_inputs,squeeze,duplexing}
 * (hasher_chip.rs:51-120).
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t state[12]; uint64_t in[8]; int n_in; uint64_t out[8]; int n_out; } challenger;
static void ch_init(challenger *c) { memset(c, 0, sizeof *c); }
static void ch_duplex(challenger *c, int len) {                 /* :107-120 */
    for (int i = 0; i < len; i++) c->state[i] = c->in[i];
    permute_cur(c->state);
    memcpy(c->out, c->state, 64); c->n_out = 8;
}
static void ch_observe(challenger *c, uint64_t v) {             /* :51-59 update: clears output buffer */
    c->n_out = 0;
    c->in[c->n_in++] = v;
    if (c->n_in == 8) { ch_duplex(c, 8); c->n_in = 0; }        /* eager chunking == chunks(RATE) at :61-72 */
}
static uint64_t ch_squeeze(challenger *c) {                     /* :73-89 */
    if (c->n_in) { ch_duplex(c, c->n_in); c->n_in = 0; }
    if (c->n_out == 0) { permute_cur(c->state); memcpy(c->out, c->state, 64); c->n_out = 8; }
    return c->out[--c->n_out];                                  /* pops from the END (:84-86) */
}

/* NOTE on eager chunking: the reference buffers all updates and absorbs them in chunks(8) at the
 * next squeeze; a full chunk duplexed early gives the same state, but `update` must also clear the
 * output buffer (it does, above) and a duplex sets the buffer -- the buffer is cleared again by the
 * next observe or consumed by the squeeze exactly as in the lazy version. */

void orc_fri_challenges(const orc_shape *s, uint64_t *rec, const uint64_t circuit_digest[4],
                        const uint64_t pi_hash[4], uint32_t num_challenges) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return;
    cur_kind = (int)s->hash_kind;
    challenger c; ch_init(&c);
    for (int i = 0; i < 4; i++) ch_observe(&c, circuit_digest[i]);                       /* plonk_verifier_chip.rs:65-67 */
    for (int i = 0; i < 4; i++) ch_observe(&c, pi_hash[i]);                              /* :69-71 */
    const uint64_t *caps = rec + L.off_init_caps;
    for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, caps[1 * L.ncap * 4 + i]); /* wires_cap :86-90 */
    for (uint32_t i = 0; i < 2 * num_challenges; i++) (void)ch_squeeze(&c);              /* betas, gammas :91-92 */
    for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, caps[2 * L.ncap * 4 + i]); /* zs_pp cap :94-98 */
    for (uint32_t i = 0; i < num_challenges; i++) (void)ch_squeeze(&c);                  /* alphas :99 */
    for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, caps[3 * L.ncap * 4 + i]); /* quotient cap :101-105 */
    uint64_t z0 = ch_squeeze(&c), z1 = ch_squeeze(&c);                                   /* zeta :106 */
    for (uint32_t i = 0; i < 2 * L.n0; i++) ch_observe(&c, rec[L.off_open0 + i]);        /* openings :108-114 */
    for (uint32_t i = 0; i < 2 * L.n1; i++) ch_observe(&c, rec[L.off_open1 + i]);
    rec[L.off_alpha] = ch_squeeze(&c); rec[L.off_alpha + 1] = ch_squeeze(&c);            /* fri_alpha :117-118 */
    for (uint32_t st = 0; st < s->num_steps; st++) {                                      /* :121-129 */
        const uint64_t *cap = rec + L.off_step_caps + (size_t)st * L.ncap * 4;
        for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, cap[i]);
        rec[L.off_betas + 2 * st] = ch_squeeze(&c); rec[L.off_betas + 2 * st + 1] = ch_squeeze(&c);
    }
    for (uint32_t i = 0; i < 2 * s->final_poly_len; i++) ch_observe(&c, rec[L.off_final_poly + i]); /* :131-135 */
    ch_observe(&c, rec[L.off_pow_witness]);                                               /* :137 */
    rec[L.off_pow_response] = ch_squeeze(&c);                                             /* :138 */
    for (uint32_t i = 0; i < s->num_query_rounds; i++) rec[L.off_indices + i] = ch_squeeze(&c); /* :140-141 */
    rec[L.off_zeta + 0] = z0; rec[L.off_zeta + 1] = z1;
    /* zeta_next = g * zeta, g = 7^((p-1)/2^degree_bits) (plonk_verifier_chip.rs:219-222) */
    uint64_t g = orc_pow(7, (ORC_P - 1) >> s->degree_bits);
    rec[L.off_zeta_next + 0] = orc_mul(z0, g); rec[L.off_zeta_next + 1] = orc_mul(z1, g);
}

/* plonk_betas, plonk_gammas, plonk_alphas (plonk_verifier_chip.rs:86-99): the prefix of orc_fri_challenges */
void orc_plonk_challenges(const orc_shape *s, const uint64_t *rec, const uint64_t circuit_digest[4],
                          const uint64_t pi_hash[4], uint32_t num_challenges, uint64_t *out) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return;
    cur_kind = (int)s->hash_kind;
    challenger c; ch_init(&c);
    for (int i = 0; i < 4; i++) ch_observe(&c, circuit_digest[i]);
    for (int i = 0; i < 4; i++) ch_observe(&c, pi_hash[i]);
    const uint64_t *caps = rec + L.off_init_caps;
    for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, caps[1 * L.ncap * 4 + i]);
    for (uint32_t i = 0; i < 2 * num_challenges; i++) out[i] = ch_squeeze(&c);
    for (uint32_t i = 0; i < L.ncap * 4; i++) ch_observe(&c, caps[2 * L.ncap * 4 + i]);
    for (uint32_t i = 0; i < num_challenges; i++) out[2 * num_challenges + i] = ch_squeeze(&c);
}
