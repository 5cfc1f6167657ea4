// Host-side (CPU) parts of the library: the Fiat-Shamir transcript that stays on the host per the
// north star, and a synthetic proof generator that stands in for plonky2's prover.
//
//  * Challenger / sv_fri_challenges  -- replaces PlonkVerifierChip::get_challenges
//    (chip/plonk/plonk_verifier_chip.rs:55-154) over TranscriptChip (chip/transcript_chip.rs) and the
//    duplex sponge HasherChip::{update,absorb_buffered_inputs,squeeze,duplexing}
//    (chip/hasher_chip.rs:51-120): rate 8, overwrite mode, squeeze pops from the END of state[0..8].
//  * sv_synth_proofs -- the reference ships no proofs and its prover (plonky2, Rust) cannot run here,
//    so valid plonky2-SHAPED proofs are built from scratch with the conventions the verifier implies:
//    LDE point of leaf i is 7 * omega^{bitrev(i)} (fri_chip.rs:152-166,262-264), salted leaves carry 4
//    extra limbs at the end (types/assigned.rs:57-71), step-tree leaf k holds the 2^arity_bits values of
//    coset k in leaf order (fri_chip.rs:279-311), fold = fri_chip.rs:168-226 generalised to 2^k (fri_fold.cuh),
//    final polynomial of 2^(degree_bits - sum arity_bits) coefficients, proof-of-work on the top bits of the squeezed response
//    (fri_chip.rs:364-376).  Trace columns are K-sparse polynomials over a shared set of degrees
//    (all < 2^degree_bits), which makes the LDE and the DEEP quotient cheap to evaluate point-wise
//    without an NTT; the verifier's work does not depend on how the columns were chosen.
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"
#include "layout.hpp"
#include "fri_fold.cuh"
#include "host_util.hpp"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace svb {

// ---------------------------------------------------------------------------------------------
struct Challenger {
    u32 kind;
    u64 state[12];
    u64 in[8];
    int n_in;
    u64 out[8];
    int n_out;
    explicit Challenger(u32 hash_kind) { memset(this, 0, sizeof *this); kind = hash_kind; }
    void duplex(int len) {
        for (int i = 0; i < len; i++) state[i] = in[i];
        permute_kind(kind, state);
        memcpy(out, state, 64);
        n_out = 8;
        n_in = 0;
    }
    void observe(u64 v) {
        n_out = 0;  // update() clears the output buffer (hasher_chip.rs:56)
        in[n_in++] = v;
        if (n_in == 8) duplex(8);  // same state as absorbing chunks(8) lazily at the next squeeze
    }
    void observe_n(const u64* v, size_t n) { for (size_t i = 0; i < n; i++) observe(v[i]); }
    u64 squeeze() {
        if (n_in) duplex(n_in);
        if (n_out == 0) {
            permute_kind(kind, state);
            memcpy(out, state, 64);
            n_out = 8;
        }
        return out[--n_out];
    }
    fp2 squeeze2() { u64 a = squeeze(); u64 b = squeeze(); return mk2(a, b); }
};

static void hash_or_noop(u32 kind, const u64* in, size_t n, u64 out[4]) {
    if (n <= 4) {
        memset(out, 0, 32);
        memcpy(out, in, n * 8);
    } else
        hash_no_pad(kind, in, n, out);
}
static void two_to_one(u32 kind, const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 st[12] = {0};
    memcpy(st, l, 32);
    memcpy(st + 4, r, 32);
    permute_kind(kind, st);
    memcpy(out, st, 32);
}

// SplitMix64 -> uniform canonical field elements by rejection
struct Rng {
    u64 s;
    explicit Rng(u64 seed) : s(seed) {}
    u64 next() {
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    u64 fe() { for (;;) { u64 v = next(); if (v < GL_P) return v; } }
};

static inline u32 bitrev(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// A Merkle tree stored as digest layers: layer 0 = leaf digests (n), ..., last layer = cap (ncap).
struct Tree {
    std::vector<std::vector<u64>> layers;  // each 4*count words
    void build(u32 kind, std::vector<u64>&& leaf_digests, u32 cap_height, int nthreads) {
        layers.clear();
        layers.push_back(std::move(leaf_digests));
        while (layers.back().size() / 4 > ((size_t)1 << cap_height)) {
            const std::vector<u64>& cur = layers.back();
            size_t n = cur.size() / 8;
            std::vector<u64> nxt(n * 4);
            parallel_for(n, n >= 256 ? nthreads : 1, [&](size_t b, size_t e) {
                for (size_t i = b; i < e; i++) two_to_one(kind, &cur[8 * i], &cur[8 * i + 4], &nxt[4 * i]);
            });
            layers.push_back(std::move(nxt));
        }
    }
    const u64* cap() const { return layers.back().data(); }
    // siblings bottom-up for leaf `index`: depth = layers.size() - 1
    void path(size_t index, u64* out) const {
        for (size_t l = 0; l + 1 < layers.size(); l++) {
            memcpy(out + 4 * l, &layers[l][4 * (index ^ 1)], 32);
            index >>= 1;
        }
    }
};

struct Circuit {
    sv_fri_shape shape;
    sv_fri_layout L;
    static const int K = 6;        // monomials per column
    u32 deg[K];                    // shared degrees, < 2^degree_bits
    u64 deg_scale[K];              // 7^deg
    std::vector<u64> coeff[4];     // [oracle][col*K + t], canonical
    std::vector<u64> wtab;         // omega^m, m < N
    u64 salt_key;
    Tree trees[4];
    u64 circuit_digest[4];
    u32 N;

    // monomial values at leaf i: m_t = 7^deg_t * omega^{bitrev(i) * deg_t}
    void monomials(u32 leaf, u64 m[K]) const {
        u32 e = bitrev(leaf, L.lde_bits);
        for (int t = 0; t < K; t++) m[t] = mulc(deg_scale[t], wtab[((u64)e * deg[t]) & (N - 1)]);
    }
    void leaf_values(int k, u32 leaf, u64* out) const {
        u64 m[K];
        monomials(leaf, m);
        u32 np = shape.oracle_num_polys[k];
        for (u32 j = 0; j < np; j++) {
            acc160 a = {0, 0, 0};
            for (int t = 0; t < K; t++) acc_mul(a, coeff[k][j * K + t], m[t]);
            out[j] = canon(acc_reduce(a));
        }
        if (L.leaf_len[k] > np) {  // 4 salt limbs at the end of a blinded leaf
            Rng r(salt_key ^ ((u64)k << 56) ^ ((u64)leaf * 0xD1342543DE82EF95ull));
            for (u32 j = np; j < L.leaf_len[k]; j++) out[j] = r.fe();
        }
    }
    // everything drawn from the seed (degrees, coefficients, salt key, circuit digest); no trees yet
    void init_params(const sv_fri_shape& s, u64 seed) {
        shape = s;
        make_layout(s, L);
        N = 1u << L.lde_bits;
        Rng rng(seed);
        u32 n = 1u << s.degree_bits;
        deg[0] = 0;
        deg[1] = 1;
        deg[2] = n - 1;
        for (int t = 3; t < K; t++) deg[t] = (u32)(rng.next() % n);
        for (int t = 0; t < K; t++) deg_scale[t] = pow(7, deg[t]);
        for (int k = 0; k < 4; k++) {
            coeff[k].resize((size_t)s.oracle_num_polys[k] * K);
            for (auto& c : coeff[k]) c = rng.fe();
        }
        salt_key = rng.next();
        for (int i = 0; i < 4; i++) circuit_digest[i] = rng.fe();
    }
    void build(const sv_fri_shape& s, u64 seed, int nthreads) {
        init_params(s, seed);
        u64 omega = pow(7, (GL_P - 1) >> L.lde_bits);
        wtab.resize(N);
        wtab[0] = 1;
        for (u32 i = 1; i < N; i++) wtab[i] = mulc(wtab[i - 1], omega);
        for (int k = 0; k < 4; k++) {
            std::vector<u64> dig((size_t)N * 4);
            u32 len = L.leaf_len[k];
            parallel_for(N, nthreads, [&](size_t b, size_t e) {
                std::vector<u64> leaf(len);
                for (size_t i = b; i < e; i++) {
                    leaf_values(k, (u32)i, leaf.data());
                    hash_or_noop(s.hash_kind, leaf.data(), len, &dig[4 * i]);
                }
            });
            trees[k].build(s.hash_kind, std::move(dig), s.cap_height, nthreads);
        }
    }
};

static fp2 pow2(fp2 b, u64 e) {
    fp2 r = mk2(1, 0);
    while (e) {
        if (e & 1) r = mul2(r, b);
        b = mul2(b, b);
        e >>= 1;
    }
    return r;
}

// One proof on a committed circuit.  `pi_hash` individualises the transcript.
static int prove(const Circuit& C, const u64 pi_hash[4], u32 num_challenges, u64* rec, int nthreads) {
    const sv_fri_shape& s = C.shape;
    const sv_fri_layout& L = C.L;
    const int K = Circuit::K;
    const u32 N = C.N;
    memset(rec, 0, (size_t)L.record_words * 8);
    for (int k = 0; k < 4; k++) memcpy(rec + L.off_init_caps + (size_t)k * L.ncap * 4, C.trees[k].cap(), (size_t)L.ncap * 32);

    Challenger ch(s.hash_kind);
    ch.observe_n(C.circuit_digest, 4);
    ch.observe_n(pi_hash, 4);
    ch.observe_n(C.trees[1].cap(), L.ncap * 4);
    for (u32 i = 0; i < 2 * num_challenges; i++) (void)ch.squeeze();  // plonk betas, gammas
    ch.observe_n(C.trees[2].cap(), L.ncap * 4);
    for (u32 i = 0; i < num_challenges; i++) (void)ch.squeeze();      // plonk alphas
    ch.observe_n(C.trees[3].cap(), L.ncap * 4);
    fp2 zeta = ch.squeeze2();
    u64 g = pow(7, (GL_P - 1) >> s.degree_bits);
    fp2 zeta_next = scale2(zeta, g);
    rec[L.off_zeta] = zeta.c0; rec[L.off_zeta + 1] = zeta.c1;
    rec[L.off_zeta_next] = zeta_next.c0; rec[L.off_zeta_next + 1] = zeta_next.c1;

    // openings: col(z) = sum_t coeff_t z^deg_t
    fp2 zp[K], znp[K];
    for (int t = 0; t < K; t++) { zp[t] = pow2(zeta, C.deg[t]); znp[t] = pow2(zeta_next, C.deg[t]); }
    auto open_at = [&](int k, u32 j, const fp2* pw) {
        fp2 acc = mk2(0, 0);
        for (int t = 0; t < K; t++) acc = add2(acc, scale2(pw[t], C.coeff[k][j * K + t]));
        return acc;
    };
    {
        u32 o = 0;
        for (int k = 0; k < 4; k++)
            for (u32 j = 0; j < s.oracle_num_polys[k]; j++, o++) {
                fp2 v = open_at(k, j, zp);
                rec[L.off_open0 + 2 * o] = v.c0; rec[L.off_open0 + 2 * o + 1] = v.c1;
            }
        for (u32 j = 0; j < s.num_zs; j++) {
            fp2 v = open_at(2, j, znp);
            rec[L.off_open1 + 2 * j] = v.c0; rec[L.off_open1 + 2 * j + 1] = v.c1;
        }
    }
    ch.observe_n(rec + L.off_open0, 2 * L.n0);
    ch.observe_n(rec + L.off_open1, 2 * L.n1);
    fp2 alpha = ch.squeeze2();
    rec[L.off_alpha] = alpha.c0; rec[L.off_alpha + 1] = alpha.c1;

    // Batched polynomials are K-sparse too: r_b(x) = sum_t (sum_i alpha^i coeff_{i,t}) x^deg_t.
    fp2 bc0[K], bc1[K], ro0 = mk2(0, 0), ro1 = mk2(0, 0);
    for (int t = 0; t < K; t++) bc0[t] = bc1[t] = mk2(0, 0);
    {
        fp2 ap = mk2(1, 0);
        u32 o = 0;
        for (int k = 0; k < 4; k++)
            for (u32 j = 0; j < s.oracle_num_polys[k]; j++, o++) {
                for (int t = 0; t < K; t++) bc0[t] = add2(bc0[t], scale2(ap, C.coeff[k][j * K + t]));
                ro0 = add2(ro0, mul2(ap, mk2(rec[L.off_open0 + 2 * o], rec[L.off_open0 + 2 * o + 1])));
                ap = mul2(ap, alpha);
            }
        ap = mk2(1, 0);
        for (u32 j = 0; j < s.num_zs; j++) {
            for (int t = 0; t < K; t++) bc1[t] = add2(bc1[t], scale2(ap, C.coeff[2][j * K + t]));
            ro1 = add2(ro1, mul2(ap, mk2(rec[L.off_open1 + 2 * j], rec[L.off_open1 + 2 * j + 1])));
            ap = mul2(ap, alpha);
        }
    }
    fp2 alpha_n1 = pow2(alpha, L.n1);

    // values of the DEEP quotient on the LDE domain, index i <-> point 7*omega^{bitrev(i)}
    std::vector<fp2> v(N);
    std::atomic<int> bad(0);
    parallel_for(N, nthreads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++) {
            u64 m[K];
            C.monomials((u32)i, m);
            fp2 r0 = mk2(0, 0), r1 = mk2(0, 0);
            for (int t = 0; t < K; t++) { r0 = add2(r0, scale2(bc0[t], m[t])); r1 = add2(r1, scale2(bc1[t], m[t])); }
            u64 x = mulc(7, C.wtab[bitrev((u32)i, L.lde_bits)]);
            fp2 d0 = sub2(mk2(x, 0), zeta), d1 = sub2(mk2(x, 0), zeta_next);
            if (is_zero2(d0) || is_zero2(d1)) { bad = 1; continue; }
            fp2 q0 = mul2(sub2(r0, ro0), inv2(d0));
            fp2 q1 = mul2(sub2(r1, ro1), inv2(d1));
            v[i] = add2(mul2(q0, alpha_n1), q1);
        }
    });
    if (bad) return -2;

    // commit phase: layer st holds the values of the current polynomial in leaf order (index i <-> point
    // shift * omega_cur^bitrev(i)); leaf k of its tree = the 2^ab values of coset k (flattened Fp2 limbs: their own
    // digest when 4 words, hashed otherwise); the next layer is the 2^ab-ary fold of every coset (fri_fold.cuh)
    std::vector<Tree> step_trees(s.num_steps);
    std::vector<std::vector<fp2>> step_vals(s.num_steps);
    u64 shift = 7;                       // coset shift of the current domain
    u32 bits = L.lde_bits;
    for (u32 st = 0; st < s.num_steps; st++) {
        const u32 ab = s.reduction_arity_bits[st], arity = 1u << ab;
        size_t cosets = v.size() >> ab;
        std::vector<u64> dig(cosets * 4);
        parallel_for(cosets, cosets >= 4096 ? nthreads : 1, [&](size_t b, size_t e) {
            u64 leaf[32];
            for (size_t k = b; k < e; k++) {
                for (u32 t = 0; t < arity; t++) { leaf[2 * t] = v[k * arity + t].c0; leaf[2 * t + 1] = v[k * arity + t].c1; }
                hash_or_noop(s.hash_kind, leaf, 2 * arity, &dig[4 * k]);
            }
        });
        step_trees[st].build(s.hash_kind, std::move(dig), s.cap_height, nthreads);
        memcpy(rec + L.off_step_caps + (size_t)st * L.ncap * 4, step_trees[st].cap(), (size_t)L.ncap * 32);
        ch.observe_n(step_trees[st].cap(), L.ncap * 4);
        fp2 beta = ch.squeeze2();
        rec[L.off_betas + 2 * st] = beta.c0; rec[L.off_betas + 2 * st + 1] = beta.c1;
        std::vector<fp2> nv(cosets);
        u32 stride = N >> bits;  // omega_cur = omega^stride
        parallel_for(cosets, cosets >= 4096 ? nthreads : 1, [&](size_t b, size_t e) {
            u64 leaf[32];
            for (size_t k = b; k < e; k++) {
                // the coset's first entry (leaf index k * arity, x_index_within_coset = 0) sits at x
                u64 x = mulc(shift, C.wtab[((u64)bitrev((u32)(k * arity), bits) * stride) & (N - 1)]);
                for (u32 t = 0; t < arity; t++) { leaf[2 * t] = v[k * arity + t].c0; leaf[2 * t + 1] = v[k * arity + t].c1; }
                nv[k] = fri_fold(ab, leaf, x, inv(x), 0, beta);
            }
        });
        step_vals[st] = std::move(v);
        v = std::move(nv);
        for (u32 t = 0; t < ab; t++) shift = mulc(shift, shift);
        bits -= ab;
    }
    // final polynomial: v holds M = 2^bits values at y_k = shift * w^{bitrev(k)}, w = omega^(N/M).
    {
        size_t M = v.size();
        u32 stride = N >> bits;
        u64 minv = inv((u64)M % GL_P);
        u64 sinv = inv(shift);
        u64 sp = 1;  // shift^-j
        for (size_t j = 0; j < M; j++) {
            fp2 acc = mk2(0, 0);
            for (size_t k = 0; k < M; k++) {
                // w^{-j*bitrev(k)} = wtab[(N - (j*bitrev(k)*stride mod N)) mod N]
                u64 e = ((u64)j * bitrev((u32)k, bits) * stride) & (N - 1);
                u64 wi = C.wtab[(N - e) & (N - 1)];
                acc = add2(acc, scale2(v[k], wi));
            }
            acc = scale2(acc, mulc(minv, sp));
            if (j < s.final_poly_len) {
                rec[L.off_final_poly + 2 * j] = acc.c0; rec[L.off_final_poly + 2 * j + 1] = acc.c1;
            } else if (!is_zero2(acc)) {
                return -3;  // degree bound violated: generator bug
            }
            sp = mulc(sp, sinv);
        }
    }
    ch.observe_n(rec + L.off_final_poly, 2 * s.final_poly_len);
    // proof of work: smallest witness whose response has proof_of_work_bits leading zero bits
    {
        u64 w = 0;
        for (;; w++) {
            Challenger c2 = ch;
            c2.observe(w);
            u64 r = c2.squeeze();
            if (s.proof_of_work_bits == 0 || (r >> (64 - s.proof_of_work_bits)) == 0) break;
        }
        rec[L.off_pow_witness] = w;
        ch.observe(w);
        rec[L.off_pow_response] = ch.squeeze();
    }
    for (u32 q = 0; q < s.num_query_rounds; q++) rec[L.off_indices + q] = ch.squeeze();

    // query rounds
    for (u32 q = 0; q < s.num_query_rounds; q++) {
        u64* qp = rec + L.header_words + (size_t)q * L.query_words;
        u32 idx = (u32)(rec[L.off_indices + q] & (N - 1));
        for (int k = 0; k < 4; k++) {
            C.leaf_values(k, idx, qp + L.q_off_init_evals[k]);
            C.trees[k].path(idx, qp + L.q_off_init_sibs[k]);
        }
        u32 cur = idx;
        for (u32 st = 0; st < s.num_steps; st++) {
            const u32 ab = s.reduction_arity_bits[st], arity = 1u << ab;
            u32 coset = cur >> ab;
            const std::vector<fp2>& sv = step_vals[st];
            u64* ev = qp + L.q_off_step_evals[st];
            for (u32 t = 0; t < arity; t++) { ev[2 * t] = sv[(size_t)coset * arity + t].c0; ev[2 * t + 1] = sv[(size_t)coset * arity + t].c1; }
            step_trees[st].path(coset, qp + L.q_off_step_sibs[st]);
            cur = coset;
        }
    }
    return 0;
}

}  // namespace svb

using namespace svb;

extern "C" int sv_fri_layout_make(const sv_fri_shape* shape, sv_fri_layout* out) {
    if (!shape || !out) return -1;
    return make_layout(*shape, *out);
}

extern "C" int sv_fri_challenges(const sv_fri_shape* shape, uint64_t* rec, const uint64_t circuit_digest[4],
                                 const uint64_t pi_hash[4], uint32_t num_challenges) {
    sv_fri_layout L;
    if (!shape || make_layout(*shape, L)) return -1;
    if (shape->hash_kind > SV_HASH_POSEIDON_BN254) return -2;
    Challenger ch(shape->hash_kind);
    ch.observe_n(circuit_digest, 4);
    ch.observe_n(pi_hash, 4);
    const u64* caps = rec + L.off_init_caps;
    ch.observe_n(caps + 1 * L.ncap * 4, L.ncap * 4);
    for (u32 i = 0; i < 2 * num_challenges; i++) (void)ch.squeeze();
    ch.observe_n(caps + 2 * L.ncap * 4, L.ncap * 4);
    for (u32 i = 0; i < num_challenges; i++) (void)ch.squeeze();
    ch.observe_n(caps + 3 * L.ncap * 4, L.ncap * 4);
    fp2 zeta = ch.squeeze2();
    fp2 zn = scale2(zeta, pow(7, (GL_P - 1) >> shape->degree_bits));
    rec[L.off_zeta] = zeta.c0; rec[L.off_zeta + 1] = zeta.c1;
    rec[L.off_zeta_next] = zn.c0; rec[L.off_zeta_next + 1] = zn.c1;
    ch.observe_n(rec + L.off_open0, 2 * L.n0);
    ch.observe_n(rec + L.off_open1, 2 * L.n1);
    fp2 a = ch.squeeze2();
    rec[L.off_alpha] = a.c0; rec[L.off_alpha + 1] = a.c1;
    for (u32 st = 0; st < shape->num_steps; st++) {
        ch.observe_n(rec + L.off_step_caps + (size_t)st * L.ncap * 4, L.ncap * 4);
        fp2 b = ch.squeeze2();
        rec[L.off_betas + 2 * st] = b.c0; rec[L.off_betas + 2 * st + 1] = b.c1;
    }
    ch.observe_n(rec + L.off_final_poly, 2 * shape->final_poly_len);
    ch.observe(rec[L.off_pow_witness]);
    rec[L.off_pow_response] = ch.squeeze();
    for (u32 q = 0; q < shape->num_query_rounds; q++) rec[L.off_indices + q] = ch.squeeze();
    return 0;
}

static inline u64 synth_circuit_seed(u64 seed, uint32_t c) { return seed ^ (0x5EED0000ull + c); }
static inline void synth_pi_hash(u64 seed, size_t i, u64 pi[4]) {
    Rng r(seed ^ (0xB2000000ull + i * 0x9E3779B97F4A7C15ull));
    for (int k = 0; k < 4; k++) pi[k] = r.fe();
}

// The transcript inputs sv_synth_proofs used: circuit digest of circuit c, public-input hash of proof i.
extern "C" int sv_synth_public_inputs(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                                      uint64_t* circuit_digests_out, uint64_t* pi_hashes_out) {
    sv_fri_layout L;
    if (!shape || make_layout(*shape, L)) return -1;
    if (n_circuits == 0) n_circuits = 1;
    if (n_circuits > n_proofs) n_circuits = (uint32_t)(n_proofs ? n_proofs : 1);
    if (circuit_digests_out)
        for (uint32_t c = 0; c < n_circuits; c++) {
            Circuit C;
            C.init_params(*shape, synth_circuit_seed(seed, c));
            memcpy(circuit_digests_out + 4 * c, C.circuit_digest, 32);
        }
    if (pi_hashes_out)
        for (size_t i = 0; i < n_proofs; i++) synth_pi_hash(seed, i, pi_hashes_out + 4 * i);
    return 0;
}

extern "C" int sv_synth_proofs(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                               uint32_t num_challenges, uint64_t* records_out, int nthreads) {
    return sv_synth_proofs_pi(shape, seed, n_circuits, n_proofs, num_challenges, nullptr, records_out, nthreads);
}

extern "C" int sv_synth_proofs_pi(const sv_fri_shape* shape, uint64_t seed, uint32_t n_circuits, size_t n_proofs,
                                  uint32_t num_challenges, const uint64_t* pi_hashes, uint64_t* records_out, int nthreads) {
    sv_fri_layout L;
    if (!shape || !records_out || make_layout(*shape, L)) return -1;
    if (shape->hash_kind > SV_HASH_POSEIDON_BN254) return -2;
    {
        u32 total = 0;
        for (u32 i = 0; i < shape->num_steps; i++) total += shape->reduction_arity_bits[i];
        if (total > shape->degree_bits || shape->final_poly_len != (1u << (shape->degree_bits - total))) return -3;
    }
    if (n_circuits == 0) n_circuits = 1;
    if (nthreads < 1) nthreads = 1;
    if (n_circuits > n_proofs) n_circuits = (uint32_t)(n_proofs ? n_proofs : 1);
    std::vector<Circuit> circuits(n_circuits);
    for (uint32_t c = 0; c < n_circuits; c++) circuits[c].build(*shape, synth_circuit_seed(seed, c), nthreads);
    // proofs are independent: parallelise across proofs when there are many, inside a proof otherwise
    std::atomic<int> err(0);
    bool outer = n_proofs >= (size_t)nthreads;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_proofs) break;
            u64 pi[4];
            if (pi_hashes) memcpy(pi, pi_hashes + 4 * i, 32);
            else synth_pi_hash(seed, i, pi);
            int rc = prove(circuits[i % n_circuits], pi, num_challenges, records_out + i * (size_t)L.record_words,
                           outer ? 1 : nthreads);
            if (rc) err = rc;
        }
    };
    if (outer) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(worker);
        for (auto& x : th) x.join();
    } else
        worker();
    return err;
}
