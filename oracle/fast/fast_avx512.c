/* ORACLE/FAST -- BENCH INFRASTRUCTURE ONLY: a CPU arm that does not flatter the GPU.
 *
 * oracle.c is written for clarity (one permutation at a time, `%`-free but scalar: 3.8 us per permutation and thread on the
 * bench box).  plonky2's native Poseidon is vectorised; timing the scalar restatement beside the GPU overstates every
 * GPU / CPU ratio by the difference.  This file is the same verifier with the Merkle work -- > 95 % of the time -- done EIGHT
 * hash chains at a time in AVX-512 lanes (eight query rounds of one proof walk the same tree shape), Poseidon-Goldilocks in
 * the reference's fast form (chip/plonk/gates/poseidon.rs:634-686) on 8 x u64 vectors: 64 x 64 -> 128 products from four
 * 32 x 32 vpmuludq, the 2^64 = 2^32 - 1 reduction with mask registers, the MDS layer on 32-bit halves with small constants.
 * Everything that is not a Merkle proof (range checks, proof of work, reduced openings, DEEP quotient, folds, final
 * polynomial) is oracle.c's own check_consistency, included below with its Merkle calls switched off.
 *
 * Bit-exactness: tests/test_oracle_fast.py checks the 8-lane permutation against orc_poseidon on random and corner states
 * and the accept bitmaps against orc_fri_verify_batch on valid and corrupted proofs of several shapes.  Only bench.py's
 * cpu_baseline / --impl reference legs and that test load this library; it needs AVX-512 F + DQ (the loader checks).
 */
#include "../oracle.c"
#include <immintrin.h>

typedef __m512i v8;
#define V8(x) _mm512_set1_epi64((long long)(x))
static inline v8 v_m32(void) { return V8(0xFFFFFFFFULL); }

/* (lo + 2^64 hi) mod p as a LOOSE u64 (any representative), orc_red128 without the final canonicalisation */
static inline v8 v_reduce128(v8 lo, v8 hi) {
    const v8 eps = V8(0xFFFFFFFFULL);
    v8 hh = _mm512_srli_epi64(hi, 32), hl = _mm512_and_si512(hi, eps);
    v8 t0 = _mm512_sub_epi64(lo, hh);
    __mmask8 b = _mm512_cmplt_epu64_mask(lo, hh);
    t0 = _mm512_mask_sub_epi64(t0, b, t0, eps);
    v8 t1 = _mm512_sub_epi64(_mm512_slli_epi64(hl, 32), hl);            /* hl * (2^32 - 1) */
    v8 t2 = _mm512_add_epi64(t0, t1);
    __mmask8 c = _mm512_cmplt_epu64_mask(t2, t1);
    return _mm512_mask_add_epi64(t2, c, t2, eps);
}
static inline void v_mul_wide(v8 a, v8 b, v8 *lo, v8 *hi) {
    const v8 m = v_m32();
    v8 ah = _mm512_srli_epi64(a, 32), bh = _mm512_srli_epi64(b, 32);
    v8 ll = _mm512_mul_epu32(a, b), lh = _mm512_mul_epu32(a, bh), hl = _mm512_mul_epu32(ah, b), hh = _mm512_mul_epu32(ah, bh);
    v8 t = _mm512_add_epi64(lh, _mm512_srli_epi64(ll, 32));             /* < 2^64 */
    v8 u = _mm512_add_epi64(hl, _mm512_and_si512(t, m));               /* < 2^64 */
    *lo = _mm512_or_si512(_mm512_slli_epi64(u, 32), _mm512_and_si512(ll, m));
    *hi = _mm512_add_epi64(hh, _mm512_add_epi64(_mm512_srli_epi64(t, 32), _mm512_srli_epi64(u, 32)));
}
static inline v8 v_mul(v8 a, v8 b) { v8 lo, hi; v_mul_wide(a, b, &lo, &hi); return v_reduce128(lo, hi); }
/* a (loose) + c (canonical) -> loose: after a wrap the sum is < c < p, so + eps cannot wrap again */
static inline v8 v_add_lc(v8 a, v8 c) {
    v8 s = _mm512_add_epi64(a, c);
    __mmask8 k = _mm512_cmplt_epu64_mask(s, a);
    return _mm512_mask_add_epi64(s, k, s, V8(0xFFFFFFFFULL));
}
static inline v8 v_canon(v8 a) {
    const v8 p = V8(ORC_P);
    __mmask8 k = _mm512_cmpge_epu64_mask(a, p);
    return _mm512_mask_sub_epi64(a, k, a, p);
}
static inline v8 v_sbox7(v8 x) {
    v8 x2 = v_mul(x, x), x4 = v_mul(x2, x2), x3 = v_mul(x, x2);
    return v_mul(x3, x4);
}
/* MDS row sums on 32-bit halves: coefficients < 2^6, 12 terms, so both half sums stay below 2^42 */
static inline void v_mds(v8 st[12]) {
    const v8 m = v_m32();
    v8 lo[12], hi[12], out[12];
    for (int i = 0; i < 12; i++) { lo[i] = _mm512_and_si512(st[i], m); hi[i] = _mm512_srli_epi64(st[i], 32); }
    for (int r = 0; r < 12; r++) {
        v8 al = _mm512_setzero_si512(), ah = _mm512_setzero_si512();
        for (int i = 0; i < 12; i++) {
            const v8 c = V8(ORC_MDS_MATRIX_CIRC[i]);
            const int j = i + r < 12 ? i + r : i + r - 12;
            al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[j], c));
            ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[j], c));
        }
        if (ORC_MDS_MATRIX_DIAG[r]) {
            const v8 d = V8(ORC_MDS_MATRIX_DIAG[r]);
            al = _mm512_add_epi64(al, _mm512_mul_epu32(lo[r], d));
            ah = _mm512_add_epi64(ah, _mm512_mul_epu32(hi[r], d));
        }
        /* value = al + 2^32 ah: low 64 bits and the carry-out word */
        v8 low = _mm512_add_epi64(al, _mm512_slli_epi64(ah, 32));
        __mmask8 k = _mm512_cmplt_epu64_mask(low, al);
        v8 high = _mm512_mask_add_epi64(_mm512_srli_epi64(ah, 32), k, _mm512_srli_epi64(ah, 32), V8(1));
        out[r] = v_reduce128(low, high);
    }
    for (int r = 0; r < 12; r++) st[r] = out[r];
}

/* eight Poseidon-Goldilocks permutations, fast form (gates/poseidon.rs:634-686); canonical in, canonical out */
static void poseidon8(v8 st[12]) {
    int rc = 0;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) st[i] = v_sbox7(v_add_lc(st[i], V8(ORC_ALL_ROUND_CONSTANTS[i + 12 * rc])));
        v_mds(st);
    }
    for (int i = 0; i < 12; i++) st[i] = v_add_lc(st[i], V8(ORC_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
    {   /* mds_partial_layer_init :504-537 */
        v8 res[12];
        res[0] = st[0];
        for (int c = 1; c < 12; c++) {
            v8 acc = _mm512_setzero_si512();
            for (int r = 1; r < 12; r++)
                acc = v_add_lc(v_mul(st[r], V8(ORC_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (c - 1)])), v_canon(acc));
            res[c] = acc;
        }
        for (int i = 0; i < 12; i++) st[i] = res[i];
    }
    for (int r = 0; r < 22; r++) {
        v8 s0 = v_sbox7(st[0]);
        if (r != 21) s0 = v_add_lc(s0, V8(ORC_FAST_PARTIAL_ROUND_CONSTANTS[r]));
        /* d = 25 s0 + sum w_hat_i st_i;  st_i += v_i s0   (:539-589) */
        v8 d = v_mul(s0, V8(ORC_MDS_MATRIX_CIRC[0] + ORC_MDS_MATRIX_DIAG[0]));
        for (int i = 1; i < 12; i++) {
            d = v_add_lc(v_mul(st[i], V8(ORC_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1])), v_canon(d));
            st[i] = v_add_lc(v_mul(s0, V8(ORC_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1])), v_canon(st[i]));
        }
        st[0] = d;
    }
    rc += 22;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) st[i] = v_sbox7(v_add_lc(st[i], V8(ORC_ALL_ROUND_CONSTANTS[i + 12 * rc])));
        v_mds(st);
    }
    for (int i = 0; i < 12; i++) st[i] = v_canon(st[i]);
}

/* states: 8 x 12 words, row-major (state of lane j at states[12 j ..]) -- the test entry point */
void orc_fast_poseidon8(uint64_t *states) {
    v8 st[12];
    uint64_t tmp[8];
    for (int i = 0; i < 12; i++) {
        for (int j = 0; j < 8; j++) tmp[j] = states[12 * j + i] % ORC_P;
        st[i] = _mm512_loadu_si512(tmp);
    }
    poseidon8(st);
    for (int i = 0; i < 12; i++) {
        _mm512_storeu_si512(tmp, st[i]);
        for (int j = 0; j < 8; j++) states[12 * j + i] = tmp[j];
    }
}

/* Eight Merkle proofs of one shape at once (merkle_proof_chip.rs:39-87 per lane): leaf[j], index[j], sibs[j], cap entry[j].
 * Returns a bit mask of the lanes whose proof holds. */
static unsigned merkle8(const uint64_t *const leaf[8], size_t leaf_len, const uint64_t idx[8], const uint64_t *const sibs[8], size_t depth,
                        const uint64_t *const cap_entry[8]) {
    v8 st[12];
    uint64_t tmp[8];
    for (int i = 0; i < 12; i++) st[i] = _mm512_setzero_si512();
    if (leaf_len <= 4) {
        for (size_t i = 0; i < leaf_len; i++) {
            for (int j = 0; j < 8; j++) tmp[j] = leaf[j][i];
            st[i] = _mm512_loadu_si512(tmp);
        }
    } else {
        for (size_t off = 0; off < leaf_len; off += 8) {
            size_t len = leaf_len - off < 8 ? leaf_len - off : 8;
            for (size_t i = 0; i < len; i++) {
                for (int j = 0; j < 8; j++) tmp[j] = leaf[j][off + i];
                st[i] = _mm512_loadu_si512(tmp);
            }
            poseidon8(st);
        }
        for (int i = 4; i < 12; i++) st[i] = _mm512_setzero_si512();
    }
    for (size_t lvl = 0; lvl < depth; lvl++) {
        __mmask8 bit = 0;
        for (int j = 0; j < 8; j++) bit |= (__mmask8)(((idx[j] >> lvl) & 1) << j);
        for (int i = 0; i < 4; i++) {
            for (int j = 0; j < 8; j++) tmp[j] = sibs[j][4 * lvl + i];
            v8 sib = _mm512_loadu_si512(tmp), cur = st[i];
            st[i] = _mm512_mask_blend_epi64(bit, cur, sib);       /* bit ? sibling : state */
            st[i + 4] = _mm512_mask_blend_epi64(bit, sib, cur);   /* bit ? state : sibling */
        }
        for (int i = 8; i < 12; i++) st[i] = _mm512_setzero_si512();
        poseidon8(st);
    }
    unsigned ok = 0xFF;
    for (int i = 0; i < 4; i++) {
        _mm512_storeu_si512(tmp, st[i]);
        for (int j = 0; j < 8; j++)
            if (tmp[j] != cap_entry[j][i]) ok &= ~(1u << j);
    }
    return ok;
}

/* verify_fri_proof with the Merkle proofs of 8 query rounds per vector; same verdict as orc_fri_verify */
int orc_fast_fri_verify(const orc_shape *s, const uint64_t *rec) {
    orc_layout L;
    if (orc_make_layout(s, &L) || s->hash_kind != 0) return orc_fri_verify(s, rec, NULL, NULL);   /* hash family B: scalar path */
    /* everything except the Merkle proofs, by the oracle's own code */
    orc_merkle_elsewhere = 1;
    int ok = orc_fri_verify(s, rec, NULL, NULL);
    orc_merkle_elsewhere = 0;
    if (!ok) return 0;            /* (a non-canonical word is caught here too, so the lanes below see canonical data) */
    const uint32_t Q = s->num_query_rounds, lde = L.lde_bits;
    for (uint32_t q0 = 0; q0 < Q; q0 += 8) {
        const uint64_t *leaf[8], *sibs[8], *cap[8];
        uint64_t idx[8];
        uint32_t qs[8];
        for (int j = 0; j < 8; j++) qs[j] = q0 + j < Q ? q0 + j : Q - 1;          /* ragged tail: repeat the last round */
        for (int k = 0; k < 4; k++) {
            for (int j = 0; j < 8; j++) {
                const uint64_t *qp = rec + L.header_words + (size_t)qs[j] * L.query_words;
                uint64_t x = rec[L.off_indices + qs[j]] & (((uint64_t)1 << lde) - 1);
                leaf[j] = qp + L.q_off_init_evals[k];
                sibs[j] = qp + L.q_off_init_sibs[k];
                idx[j] = x;
                cap[j] = rec + L.off_init_caps + ((size_t)k * L.ncap + (size_t)(x >> (lde - s->cap_height))) * 4;
            }
            if (merkle8(leaf, L.leaf_len[k], idx, sibs, L.init_depth, cap) != 0xFF) return 0;
        }
        uint32_t shift = 0;
        for (uint32_t i = 0; i < s->num_steps; i++) {
            shift += s->reduction_arity_bits[i];
            for (int j = 0; j < 8; j++) {
                const uint64_t *qp = rec + L.header_words + (size_t)qs[j] * L.query_words;
                uint64_t x = rec[L.off_indices + qs[j]] & (((uint64_t)1 << lde) - 1);
                leaf[j] = qp + L.q_off_step_evals[i];
                sibs[j] = qp + L.q_off_step_sibs[i];
                idx[j] = x >> shift;
                cap[j] = rec + L.off_step_caps + ((size_t)i * L.ncap + (size_t)(x >> (lde - s->cap_height))) * 4;
            }
            if (merkle8(leaf, (size_t)2 << s->reduction_arity_bits[i], idx, sibs, L.step_depth[i], cap) != 0xFF) return 0;
        }
    }
    return 1;
}

typedef struct { const orc_shape *s; const uint64_t *records; size_t n, stride; uint8_t *ok; int tid, nt; } fast_arg;
static void *fast_worker(void *p) {
    fast_arg *a = (fast_arg *)p;
    for (size_t i = (size_t)a->tid; i < a->n; i += (size_t)a->nt) a->ok[i] = (uint8_t)orc_fast_fri_verify(a->s, a->records + i * a->stride);
    return NULL;
}
void orc_fast_fri_verify_batch(const orc_shape *s, const uint64_t *records, size_t n, uint32_t *bitmap, int nthreads) {
    orc_layout L;
    if (orc_make_layout(s, &L)) return;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    uint8_t *ok = (uint8_t *)calloc(n ? n : 1, 1);
    pthread_t th[256];
    fast_arg args[256];
    for (int t = 0; t < nthreads; t++) {
        args[t] = (fast_arg){s, records, n, L.record_words, ok, t, nthreads};
        pthread_create(&th[t], NULL, fast_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    for (size_t w = 0; w < (n + 31) / 32; w++) {
        uint32_t word = 0;
        for (size_t i = w * 32; i < n && i < w * 32 + 32; i++) word |= (uint32_t)ok[i] << (i & 31);
        bitmap[w] = word;
    }
    free(ok);
}
